"""Lock-step evaluation of many sequential optimisers.

SciPy's Powell / Nelder-Mead are sequential per start.  The reference runs its restarts one after
another (gpUtils.py:223-247, utility.py:332-366), each objective call costing one GP factorisation
or one predict.  Here every restart runs in its own thread with an unmodified SciPy optimiser; the
objective calls of all live restarts rendezvous and are evaluated by ONE batched device launch.
Per restart the sequence of objective values -- hence the optimiser path -- is exactly what the
sequential reference loop would see.

Scheduling is baton passing: exactly one thread is runnable at any time.  A worker that reaches an
objective call parks its argument and hands the baton to the next worker that has not yet submitted
in this round; the last one evaluates the whole batch and wakes the first.  (A condition variable
with notify_all costs ~80 us per worker per round in GIL hand-offs; the baton costs ~10 us.)
"""
import threading

import numpy as np


class LockstepEvaluator(object):
    def __init__(self, nworkers, batch_fn):
        self._batch_fn = batch_fn
        self._locks = [threading.Lock() for _ in range(nworkers)]
        for lk in self._locks:
            lk.acquire()                      # every worker starts parked
        self._active = list(range(nworkers))  # ring order
        self._pending = {}
        self._results = {}
        self._error = None
        self.nbatches = 0
        self.nevals = 0

    # ---- baton logic (always called by the single running thread) ----------------------------
    def _flush(self):
        ids = sorted(self._pending)
        xs = [self._pending[i] for i in ids]
        self._pending = {}
        try:
            vals = self._batch_fn(xs)
        except BaseException as e:            # every worker re-raises it when it wakes
            self._error = e
            vals = [np.nan] * len(ids)
        self.nbatches += 1
        self.nevals += len(ids)
        for i, v in zip(ids, vals):
            self._results[i] = v

    def _pass_baton(self, after):
        """Wake the next worker: the first active one (ring order after `after`) that has not submitted
        this round; if all have, evaluate the batch first and wake the first active worker."""
        if not self._active:
            return
        n = len(self._locks)
        order = sorted(self._active, key=lambda w: (w - after - 1) % n)
        for w in order:
            if w not in self._pending:
                self._locks[w].release()
                return
        self._flush()
        self._locks[order[0]].release()

    def start(self):
        self._pass_baton(-1)

    def wait_turn(self, wid):
        self._locks[wid].acquire()
        if self._error is not None:
            self._pass_baton(wid)             # let the others wake up and fail too
            raise self._error

    def evaluate(self, wid, x):
        self._pending[wid] = np.array(x, dtype=np.float64, copy=True)
        self._pass_baton(wid)
        self._locks[wid].acquire()            # parked until this round has been evaluated
        if self._error is not None:
            if wid in self._active:
                self._active.remove(wid)
            self._pass_baton(wid)
            raise self._error
        return self._results.pop(wid)

    def finish(self, wid):
        if wid in self._active:
            self._active.remove(wid)
            self._pass_baton(wid)


def run_lockstep(nworkers, batch_fn, worker_fn):
    """Run ``worker_fn(wid, f)`` for wid in range(nworkers) in threads, where ``f(x)`` is the
    batched objective.  Returns (list of results in wid order, evaluator)."""
    ev = LockstepEvaluator(nworkers, batch_fn)
    out = [None] * nworkers
    errs = [None] * nworkers

    def body(wid):
        try:
            ev.wait_turn(wid)
            out[wid] = worker_fn(wid, lambda x: ev.evaluate(wid, x))
        except BaseException as e:
            errs[wid] = e
        finally:
            ev.finish(wid)

    threads = [threading.Thread(target=body, args=(w,), daemon=True) for w in range(nworkers)]
    for t in threads:
        t.start()
    ev.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out, ev
