"""Data-parallel plumbing (NEW capability; the reference is single-device, train.py:111,116 -- SURVEY 0.8 / 8e).

One process per GPU; the only exchange of the training step is ONE sum all-reduce of the flat fp32 gradient arena
(NCCL over NVLink/NVSwitch on GPUs, gloo in the CPU tests).  Averaging (1/world) is folded into the optimiser kernel as
`grad_scale`; clip-by-global-norm and Adam run AFTER the reduce, identically on every rank; BatchNorm statistics stay
per replica."""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment (RANK / WORLD_SIZE / MASTER_*); no-op for 1 process."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    # CRNN_DIST_BACKEND=gloo: several ranks share one GPU (NCCL refuses duplicate devices) -- used by the 2-rank test on 1-GPU boxes
    backend = backend or os.environ.get("CRNN_DIST_BACKEND") or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
    dist.init_process_group(backend, **kw)


def backend():
    return dist.get_backend() if dist.is_available() and dist.is_initialized() else None


def native_comm_for(model, fused=True):
    """Give `model`'s engine its own NCCL communicator over the ranks of the default process group (ncclUniqueId created by the C ABI on rank
    0 and broadcast through torch.distributed) and, with `fused`, let the training step issue its bucketed gradient all-reduce itself
    (include/crnn_b200.h: crnn_comm_init_rank / crnn_set_dp_fused).  Returns False when not applicable: single process, a non-NCCL process
    group (the CPU / shared-GPU test modes), or CRNN_DP_NATIVE=0 (A/B against the torch.distributed all-reduce)."""
    if world_size() <= 1 or backend() != "nccl" or os.environ.get("CRNN_DP_NATIVE") == "0":
        return False
    from . import _lib
    lib = model.lib
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank() == 0:
        _lib.check(lib.crnn_nccl_unique_id(uid.data_ptr()))
    uid_dev = uid.to(model.device)
    dist.broadcast(uid_dev, src=0)
    uid = uid_dev.cpu().contiguous()
    with torch.cuda.device(model.device):
        _lib.check(lib.crnn_comm_init_rank(model.handle, uid.data_ptr(), world_size(), rank()))
        if fused:
            _lib.check(lib.crnn_set_dp_fused(model.handle, 1))
    return True


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(flat: torch.Tensor) -> float:
    """In-place sum all-reduce of the flat gradient arena; returns the scale (1/world) the optimiser must apply."""
    w = world_size()
    if w > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    return 1.0 / w


def allreduce_mean_(t: torch.Tensor) -> torch.Tensor:
    """In-place mean over ranks of a small tensor (the monitored loss: every rank must take the same early-stopping / checkpoint decisions)."""
    w = world_size()
    if w > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        t /= w
    return t


def steps_per_epoch(n_items: int, batch_size: int, w=None) -> int:
    """Number of optimiser steps per epoch, IDENTICAL on every rank: derived from the largest shard (ceil(n/world)), never from the local one --
    ranks issuing different numbers of gradient all-reduces would hang the job (the batch generator wraps around its file list, utils.py:454-511)."""
    w = world_size() if w is None else w
    return -(-(-(-n_items // w)) // batch_size)


def broadcast_(flat: torch.Tensor, src=0):
    """Make every replica start from rank `src`'s parameters."""
    if world_size() > 1:
        dist.broadcast(flat, src=src)


def shard_batch(n_items: int, r=None, w=None):
    """Contiguous shard [lo, hi) of a global batch for this rank (text lines are independent: no data-path collective)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    per, rem = divmod(n_items, w)
    lo = r * per + min(r, rem)
    return lo, lo + per + (1 if r < rem else 0)
