"""Host plumbing of row-sharded chains (include/bayesrr_b200.h, `brr_comm`).

The library needs two host collectives at set-up time and at the run boundaries -- a sum-allreduce of doubles and an
all-gather of bytes (CUDA IPC handles of the exchange windows).  Two implementations:

  TorchComm    one process per GPU (torchrun): `torch.distributed` over a gloo group for the host buffers; the NCCL
               default group, when there is one, stays free for the caller's device collectives.
  ThreadGroup  several ranks as threads of ONE process (how a single R session would drive 8 GPUs; also lets a
               1-GPU box exercise the exchange protocol with two ranks on the same device).

Everything inside the iteration loop -- the per-block exchange of partial X_b^T eps, the Gram sum, the end-of-sweep
sums -- runs device-to-device over NVLink peer memory inside the kernels; these call-backs never see it.
"""
import ctypes as C
import threading

import numpy as np

_dp = C.POINTER(C.c_double)
ALLREDUCE_T = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int64)
ALLGATHER_T = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


class CommStruct(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("allreduce_sum", ALLREDUCE_T), ("allgather", ALLGATHER_T),
                ("ctx", C.c_void_p)]


def shard_bounds(n_rows, world):
    """rows [lo, hi) of every rank: contiguous, sizes differ by at most one 64-row unit (the workers' granularity)"""
    units = (n_rows + 63) // 64
    cuts = [min(n_rows, 64 * (units * r // world)) for r in range(world + 1)]
    cuts[-1] = n_rows
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class Comm:
    """one rank's brr_comm; keeps the ctypes call-backs alive"""

    def __init__(self, rank, world, allreduce, allgather):
        self.rank, self.world = rank, world
        self.errors = []

        def _ar(ctx, buf, n):
            try:
                allreduce(np.ctypeslib.as_array(buf, shape=(n,)))
                return 0
            except Exception as e:          # never let an exception cross the C frame
                self.errors.append(e)
                return 1

        def _ag(ctx, send, recv, nbytes):
            try:
                s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,))
                r = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(world, nbytes))
                allgather(s, r)
                return 0
            except Exception as e:
                self.errors.append(e)
                return 1
        self._cb = (ALLREDUCE_T(_ar), ALLGATHER_T(_ag))
        self.struct = CommStruct(rank, world, self._cb[0], self._cb[1], None)

    def byref(self):
        return C.byref(self.struct)

    def selftest(self, buf, token):
        from . import lib, _check
        buf = np.ascontiguousarray(buf, dtype=np.float64)
        got = np.zeros(self.world, dtype=np.int64)
        _check(lib().brr_comm_selftest(self.byref(), buf.ctypes.data_as(_dp), C.c_int64(len(buf)), C.c_int64(token),
                                       got.ctypes.data_as(C.POINTER(C.c_int64))))
        return buf, got


class ThreadGroup:
    """`world` ranks as threads of this process"""

    def __init__(self, world):
        self.world = world
        self._bar = threading.Barrier(world)
        self._slots = [None] * world

    def comm(self, rank):
        def allreduce(arr):
            self._slots[rank] = arr.copy()
            self._bar.wait()
            total = self._slots[0].copy()
            for r in range(1, self.world):          # rank order on every rank: identical bits
                total += self._slots[r]
            self._bar.wait()
            arr[:] = total

        def allgather(send, recv):
            self._slots[rank] = send.copy()
            self._bar.wait()
            for r in range(self.world):
                recv[r, :] = self._slots[r]
            self._bar.wait()
        return Comm(rank, self.world, allreduce, allgather)

    def run(self, fn):
        """fn(rank, comm) on `world` threads; returns the results in rank order, re-raises the first failure"""
        out, err = [None] * self.world, [None] * self.world

        def body(r):
            try:
                out[r] = fn(r, self.comm(r))
            except BaseException as e:
                err[r] = e
                self._bar.abort()
        ts = [threading.Thread(target=body, args=(r,)) for r in range(self.world)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, threading.BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out


def torch_comm(group=None):
    """brr_comm over torch.distributed (one process per GPU).  `group` must be a gloo group (host tensors); when None a
    gloo group over all ranks is created next to the default (NCCL) group."""
    import torch
    import torch.distributed as dist
    if group is None:
        group = dist.group.WORLD if dist.get_backend() == "gloo" else dist.new_group(backend="gloo")
    rank, world = dist.get_rank(group), dist.get_world_size(group)

    def allreduce(arr):
        t = torch.from_numpy(arr)               # shares memory with the C buffer
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

    def allgather(send, recv):
        outs = [torch.empty(len(send), dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(outs, torch.from_numpy(send.copy()), group=group)
        for r in range(world):
            recv[r, :] = outs[r].numpy()
    c = Comm(rank, world, allreduce, allgather)
    c.group = group
    return c
