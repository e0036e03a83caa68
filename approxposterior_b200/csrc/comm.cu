// Multi-GPU entry points of the C-ABI: one handle per GPU, NCCL over NVLink 5 / NVSwitch.
//
// The path shards by independent units (queries, ensembles, restarts -- SURVEY 8e), so there are exactly two
// exchanges: the factorised GP state is replicated once per GP update (apgp_comm_broadcast_factor), and results are
// concatenated once per call (apgp_comm_allgather: chains, candidate scores, restart results).  No collective sits
// inside a kernel's inner loop, hence plain NCCL collectives on the handle's stream -- there is no compute step that
// is immediately followed by a transfer of its own output to fuse with.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): libapgp.so carries no link-time dependency on it, a process that
// already loaded PyTorch's bundled NCCL shares that copy, and single-GPU users never touch it.
#include "../../include/apgp.h"
#include "apgp_internal.h"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <string>

namespace apgp {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

static NcclApi g_nccl;

const char* nccl_load_error() { return g_nccl.err.c_str(); }

NcclApi* nccl_api() {
  if (g_nccl.lib) return &g_nccl;
  const char* names[] = {getenv("APGP_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    if (!n) continue;
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) { g_nccl.err = std::string("cannot load NCCL (libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return nullptr; }
#define APGP_SYM(field, name)                                                                     \
  g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(lib, name));                      \
  if (!g_nccl.field) { g_nccl.err = std::string("NCCL symbol missing: ") + name; return nullptr; }
  APGP_SYM(GetUniqueId, "ncclGetUniqueId")
  APGP_SYM(CommInitRank, "ncclCommInitRank")
  APGP_SYM(CommDestroy, "ncclCommDestroy")
  APGP_SYM(Broadcast, "ncclBroadcast")
  APGP_SYM(AllGather, "ncclAllGather")
  APGP_SYM(GroupStart, "ncclGroupStart")
  APGP_SYM(GroupEnd, "ncclGroupEnd")
  APGP_SYM(GetErrorString, "ncclGetErrorString")
#undef APGP_SYM
  g_nccl.lib = lib;
  return &g_nccl;
}

static std::string g_comm_err;
const char* comm_last_error() { return g_comm_err.c_str(); }
static int nccl_fail(NcclApi* a, const char* what, ncclResult_t r) {
  g_comm_err = std::string(what) + ": " + (a && a->GetErrorString ? a->GetErrorString(r) : "NCCL error");
  return -1;
}
#define NCCL_TRY(call, what) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return nccl_fail(a, what, r_); } while (0)

int comm_unique_id(char out[128]) {
  NcclApi* a = nccl_api();
  if (!a) { g_comm_err = nccl_load_error(); return -1; }
  ncclUniqueId id;
  NCCL_TRY(a->GetUniqueId(&id), "ncclGetUniqueId");
  static_assert(sizeof(id) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out, &id, 128);
  return 0;
}
int comm_init(void** comm_out, const char id_bytes[128], int rank, int world) {
  NcclApi* a = nccl_api();
  if (!a) { g_comm_err = nccl_load_error(); return -1; }
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  ncclComm_t c = nullptr;
  NCCL_TRY(a->CommInitRank(&c, world, id, rank), "ncclCommInitRank");
  *comm_out = c;
  return 0;
}
int comm_destroy(void* comm) {
  NcclApi* a = nccl_api();
  if (!a || !comm) return 0;
  NCCL_TRY(a->CommDestroy(static_cast<ncclComm_t>(comm)), "ncclCommDestroy");
  return 0;
}
int comm_broadcast_bytes(void* comm, void* buf_dev, size_t bytes, int root, cudaStream_t st) {
  NcclApi* a = nccl_api();
  if (!a) { g_comm_err = nccl_load_error(); return -1; }
  NCCL_TRY(a->Broadcast(buf_dev, buf_dev, bytes, ncclChar, root, static_cast<ncclComm_t>(comm), st), "ncclBroadcast");
  return 0;
}
int comm_allgather_doubles(void* comm, const double* send_dev, double* recv_dev, size_t count, cudaStream_t st) {
  NcclApi* a = nccl_api();
  if (!a) { g_comm_err = nccl_load_error(); return -1; }
  NCCL_TRY(a->AllGather(send_dev, recv_dev, count, ncclDouble, static_cast<ncclComm_t>(comm), st), "ncclAllGather");
  return 0;
}
int comm_group_start() { NcclApi* a = nccl_api(); if (!a) { g_comm_err = nccl_load_error(); return -1; } NCCL_TRY(a->GroupStart(), "ncclGroupStart"); return 0; }
int comm_group_end() { NcclApi* a = nccl_api(); if (!a) { g_comm_err = nccl_load_error(); return -1; } NCCL_TRY(a->GroupEnd(), "ncclGroupEnd"); return 0; }

}  // namespace apgp
