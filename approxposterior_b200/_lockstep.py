"""Lock-step evaluation of many sequential optimisers.

SciPy's Powell / Nelder-Mead are sequential per start.  The reference runs its restarts one after
another (gpUtils.py:223-247, utility.py:332-366), each objective call costing one GP factorisation
or one predict.  Here every restart runs in its own thread with an unmodified SciPy optimiser; the
objective calls of all live restarts rendezvous and are evaluated by ONE batched device launch.
Per restart the sequence of objective values -- hence the optimiser path -- is exactly what the
sequential reference loop would see.
"""
import threading

import numpy as np


class LockstepEvaluator(object):
    def __init__(self, nworkers, batch_fn):
        self._batch_fn = batch_fn
        self._cv = threading.Condition()
        self._active = nworkers
        self._pending = {}
        self._results = {}
        self._error = None
        self.nbatches = 0
        self.nevals = 0

    def _flush_locked(self):
        ids = sorted(self._pending)
        xs = [self._pending[i] for i in ids]
        self._pending = {}
        try:
            vals = self._batch_fn(xs)
        except BaseException as e:      # propagate to every waiting worker
            self._error = e
            vals = [np.nan] * len(ids)
        self.nbatches += 1
        self.nevals += len(ids)
        for i, v in zip(ids, vals):
            self._results[i] = v
        self._cv.notify_all()

    def evaluate(self, wid, x):
        with self._cv:
            self._pending[wid] = np.array(x, dtype=np.float64, copy=True)
            if len(self._pending) >= self._active:
                self._flush_locked()
            while wid not in self._results:
                self._cv.wait()
            if self._error is not None:
                raise self._error
            return self._results.pop(wid)

    def finish(self, wid):
        with self._cv:
            self._active -= 1
            if self._pending and len(self._pending) >= self._active:
                self._flush_locked()


def run_lockstep(nworkers, batch_fn, worker_fn):
    """Run ``worker_fn(wid, f)`` for wid in range(nworkers) in threads, where ``f(x)`` is the
    batched objective.  Returns (list of results in wid order, evaluator)."""
    ev = LockstepEvaluator(nworkers, batch_fn)
    out = [None] * nworkers
    errs = [None] * nworkers

    def body(wid):
        try:
            out[wid] = worker_fn(wid, lambda x: ev.evaluate(wid, x))
        except BaseException as e:
            errs[wid] = e
        finally:
            ev.finish(wid)

    threads = [threading.Thread(target=body, args=(w,), daemon=True) for w in range(nworkers)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errs:
        if e is not None:
            raise e
    return out, ev
