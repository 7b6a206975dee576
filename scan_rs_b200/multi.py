"""One host thread driving every GPU of the box: mirror of sb_multi (csrc/multi.cu; SURVEY 8b "one sb_ctx drives all GPUs from
one host thread").  `MultiContext.run(fn)` executes fn(rank, ctx) on one library worker thread per GPU, concurrently; inside it
every rank makes the ordinary calls on its own cell shard and the collectives meet across the workers."""
from __future__ import annotations

import ctypes as C
import traceback
from typing import Callable, List, Optional, Sequence

from . import _lib as L
from .sqz import Context

_RANK_FN = C.CFUNCTYPE(C.c_int, C.c_int, C.c_void_p, C.c_void_p)


class MultiContext:
    def __init__(self, devices: Optional[Sequence[int]] = None, n: Optional[int] = None):
        devs = list(devices) if devices is not None else list(range(n or 1))
        arr = (C.c_int * len(devs))(*devs)
        self._h = C.c_void_p()
        L.check(L.lib().sb_multi_init(C.c_int(len(devs)), arr, C.byref(self._h)))
        self.n = len(devs)
        self.contexts: List[Context] = []
        for r in range(self.n):
            h = C.c_void_p()
            L.check(L.lib().sb_multi_ctx(self._h, C.c_int(r), C.byref(h)))
            self.contexts.append(Context._adopt(h, devs[r], self.n, r))

    def run(self, fn: Callable[[int, Context], None]) -> list:
        """fn(rank, ctx) on every rank's worker thread; returns the per-rank return values.  An exception in any rank is re-raised."""
        results, errors = [None] * self.n, [None] * self.n

        def tramp(rank, _ctx, _user):
            try:
                results[rank] = fn(rank, self.contexts[rank])
                return 0
            except L.ScanB200Error as e:
                errors[rank] = e
                return e.code or L.SB_ERR_INVALID_ARG
            except BaseException as e:  # noqa: BLE001 -- must not unwind through the C frames of the worker
                errors[rank] = e
                traceback.print_exc()
                return L.SB_ERR_INVALID_ARG

        cb = _RANK_FN(tramp)
        rc = L.lib().sb_multi_run(self._h, cb, None)
        for e in errors:
            if e is not None:
                raise e
        L.check(rc)
        return results

    def close(self):
        if self._h:
            for c in self.contexts:
                c._release_handles()
            L.lib().sb_multi_shutdown(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
