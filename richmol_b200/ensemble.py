"""Ensemble sharding across GPUs: one process per GPU, rows of `vecs` split across ranks.

The reference's only data-parallel pattern for the TDSE is user-level slicing of the initial states by
MPI rank (docs/source/notebooks/tdse_mpi.ipynb:268-272): ensemble members are independent
(richmol/tdse.py:379,399), operators are replicated, nothing crosses ranks during a step.  The one
exchange the path needs is the sum of the Boltzmann-weighted observables at an output time: a single
`all_reduce(SUM)` of a few float64 values (NCCL over NVLink on GPU ranks, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(nstates, rank, world):
    """Contiguous, balanced row range [lo, hi) of rank `rank` (first `nstates % world` ranks get one
    more row)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, extra = divmod(int(nstates), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_rows(vecs, rank=None, world=None):
    """This rank's rows of the ensemble (a view); rank/world default to the torch.distributed group."""
    if rank is None or world is None:
        import torch.distributed as dist
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(len(vecs), rank, world)
    return vecs[lo:hi]


def allreduce_sum(values, group=None):
    """Sum of a small vector of observables over all ranks; returns the same type it was given
    (numpy array / python scalar -> numpy array, torch tensor -> tensor, in place)."""
    import torch
    import torch.distributed as dist
    if isinstance(values, torch.Tensor):
        if dist.is_initialized() and dist.get_world_size(group) > 1:
            if values.is_complex():
                dist.all_reduce(torch.view_as_real(values), group=group)
            else:
                dist.all_reduce(values, group=group)
        return values
    arr = np.atleast_1d(np.asarray(values))
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return arr
    cplx = np.iscomplexobj(arr)
    buf = np.ascontiguousarray(arr, dtype=np.complex128 if cplx else np.float64)
    t = torch.from_numpy(buf.view(np.float64) if cplx else buf)
    if dist.get_backend(group) == "nccl":
        t = t.cuda()
        dist.all_reduce(t, group=group)
        t = t.cpu()
    else:
        dist.all_reduce(t, group=group)
    out = t.numpy()
    return out.view(np.complex128) if cplx else out


def ensemble_expectation(O, vecs_local, group=None):
    """sum_i <v_i|O|v_i> over the WHOLE ensemble: local fused expectation kernel + one all-reduce."""
    from .tdse import expectation
    ev = expectation(O, vecs_local)
    local = ev.sum()
    import torch
    if isinstance(local, torch.Tensor):
        return allreduce_sum(local.reshape(1), group)[0]
    return allreduce_sum(np.array([local]), group)[0]


def bind_to_gpu_numa(device_index):
    """Pins the calling process to the CPU cores of the NUMA node the GPU hangs off (NVML's ideal CPU affinity), so
    that pinned staging buffers are first-touched on that node and host <-> device copies of the per-rank shard do not
    cross the socket interconnect.  With one process per GPU this is what keeps the host-buffer path (`TDSE.update` on
    numpy arrays) from collapsing when all GPUs of a box move their shards at once.  Returns the CPU set or None."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device_index]) if vis else int(device_index)
        h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return sorted(cpus)
    except Exception:
        pass
    return None
