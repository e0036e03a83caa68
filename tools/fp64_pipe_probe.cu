// FP64 pipe probe for sm_100a: DMMA.8x8x4 peak, DFMA peak, and whether they share a pipe.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_pipe_probe tools/fp64_pipe_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));
}

// mode: 0 = all warps DMMA, 1 = all warps DFMA, 2 = even warps DMMA / odd warps DFMA
template<int NACC>
__global__ void probe(double* out, int iters, int mode){
  int warp = threadIdx.x >> 5;
  bool do_mma = (mode==0) || (mode==2 && (warp&1)==0);
  double a = 1.0 + 1e-9*threadIdx.x, b = 1.0 - 1e-9*threadIdx.x;
  double c[2*NACC];
  #pragma unroll
  for(int i=0;i<2*NACC;i++) c[i]=0.0;
  if(do_mma){
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<NACC;i++) dmma(c[2*i], c[2*i+1], a, b);
    }
  } else {
    for(int it=0; it<iters; it++){
      #pragma unroll
      for(int i=0;i<2*NACC;i++) c[i] = fma(a, b, c[i]);   // 2*NACC dependent-chain DFMAs per iter
    }
  }
  double s=0;
  #pragma unroll
  for(int i=0;i<2*NACC;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}

// layout check: C = A(8x4) * B(4x8)
__global__ void layout_check(const double* A, const double* B, double* C){
  int l = threadIdx.x;
  double a = A[(l/4)*4 + (l%4)];       // A[row=l/4][k=l%4], row-major 8x4
  double b = B[(l%4)*8 + (l/4)];       // B[k=l%4][n=l/4], row-major 4x8
  double c0=0,c1=0;
  dmma(c0,c1,a,b);
  C[(l/4)*8 + (l%4)*2 + 0] = c0;
  C[(l/4)*8 + (l%4)*2 + 1] = c1;
}

int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0));
  printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
  // layout
  {
    double hA[32], hB[32], hC[64], ref[64];
    for(int i=0;i<32;i++){hA[i]=(i*7%11)-3.0; hB[i]=(i*5%13)-4.0;}
    for(int r=0;r<8;r++)for(int n=0;n<8;n++){double s=0;for(int k=0;k<4;k++)s+=hA[r*4+k]*hB[k*8+n];ref[r*8+n]=s;}
    double *dA,*dB,*dC; CK(cudaMalloc(&dA,256));CK(cudaMalloc(&dB,256));CK(cudaMalloc(&dC,512));
    CK(cudaMemcpy(dA,hA,256,cudaMemcpyHostToDevice));CK(cudaMemcpy(dB,hB,256,cudaMemcpyHostToDevice));
    layout_check<<<1,32>>>(dA,dB,dC); CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(hC,dC,512,cudaMemcpyDeviceToHost));
    int bad=0; for(int i=0;i<64;i++) if(hC[i]!=ref[i]) bad++;
    printf("layout_check m8n8k4: %s (%d mismatches)\n", bad?"FAIL":"OK", bad);
  }
  int nsm = p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out, sizeof(double)*nsm*4*1024));
  cudaEvent_t e0,e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int iters = 20000;
  const char* names[3]={"DMMA","DFMA","MIX(even DMMA/odd DFMA)"};
  for(int mode=0; mode<3; mode++){
    for(int threads=128; threads<=1024; threads*=2){
      for(int ctas=1; ctas<=2; ctas++){
        if(threads*ctas>2048) continue;
        float best=1e30f;
        for(int rep=0;rep<4;rep++){
          CK(cudaEventRecord(e0));
          probe<8><<<nsm*ctas,threads>>>(out,iters,mode);
          CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
          float ms; CK(cudaEventElapsedTime(&ms,e0,e1)); if(ms<best)best=ms;
        }
        double warps = (double)nsm*ctas*threads/32;
        double mma_w = mode==0?warps: mode==2?warps/2:0;
        double fma_w = mode==1?warps: mode==2?warps/2:0;
        double mma_flops = mma_w*iters*8.0*256*2;
        double fma_flops = fma_w*iters*16.0*32*2;
        printf("%-26s threads=%4d ctas/SM=%d  %.3f ms  DMMA %.2f TF  DFMA %.2f TF  total %.2f TF\n", names[mode], threads, ctas, best,
               mma_flops/best*1e-9, fma_flops/best*1e-9, (mma_flops+fma_flops)/best*1e-9);
      }
    }
  }
  return 0;
}
