"""Host-side plumbing for the multi-GPU path: one process per GPU (torchrun), the voxel columns of
``Pt = A.K`` sharded contiguously over ranks.

``torch.distributed`` (gloo) is used ONLY for the control plane -- distributing the NCCL unique id,
barriers, gathering the small result vectors and max-over-ranks timings.  The one data-path
collective (all-reduce of the AkA partial sums) runs inside ``libgeobo_b200.so`` over NCCL/NVLink.
"""
import os

import numpy as np

_state = {"rank": 0, "world": 1, "initialized": False}


def rank():
    return _state["rank"]


def world_size():
    return _state["world"]


def shard_columns(n_vox, world, rank, align=128):
    """Contiguous voxel-column shard [c0, c1) of rank ``rank``; boundaries are multiples of ``align``
    (the GEMM tile width; the C ABI requires multiples of 16).  Every rank must own at least one column."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world %r/%r" % (rank, world))
    per = -(-n_vox // world)
    per = -(-per // align) * align
    c0 = rank * per
    c1 = min(n_vox, c0 + per)
    if c0 >= n_vox:
        raise ValueError("cube with %d voxels is too small to shard over %d ranks at alignment %d" % (n_vox, world, align))
    return c0, c1


def init_from_env(ctx=None, backend="gloo"):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  With
    WORLD_SIZE <= 1 this is a no-op.  If ``ctx`` (a ``_lib.Context``) is given, its NCCL communicator
    is created from a unique id broadcast by rank 0."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    _state.update(rank=rk, world=world)
    if world <= 1:
        return rk, world
    import torch.distributed as td
    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        td.init_process_group(backend=backend, rank=rk, world_size=world)
    _state["initialized"] = True
    if ctx is not None and ctx.nranks != world:
        uid = [ctx.comm_unique_id() if rk == 0 else None]
        td.broadcast_object_list(uid, src=0)
        ctx.comm_init(uid[0], rk, world)
    return rk, world


def barrier():
    if _state["world"] > 1:
        import torch.distributed as td
        td.barrier()


def max_over_ranks(value):
    if _state["world"] <= 1:
        return float(value)
    import torch
    import torch.distributed as td
    t = torch.tensor([float(value)], dtype=torch.float64)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t[0])


def allgather_columns(local, n_vox, align=128, ctx=None):
    """``local``: (k, ncol_local) array of this rank's voxel columns -> (k, n_vox) on every rank.
    With a ``_lib.Context`` that has an NCCL communicator the gather runs over NCCL/NVLink inside the library
    (``gb_comm_allgather``); otherwise over the host control plane (gloo)."""
    world, rk = _state["world"], _state["rank"]
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world <= 1:
        return local
    if ctx is not None and getattr(ctx, "nranks", 1) == world:
        k = local.shape[0]
        per = shard_columns(n_vox, world, 0, align)[1]
        buf = np.zeros((k, per))
        buf[:, :local.shape[1]] = local
        outs = ctx.allgather(buf).reshape(world, k, per)
        full = np.empty((k, n_vox))
        for r in range(world):
            c0, c1 = shard_columns(n_vox, world, r, align)
            full[:, c0:c1] = outs[r][:, :c1 - c0]
        return full
    import torch
    import torch.distributed as td
    k = local.shape[0]
    per = shard_columns(n_vox, world, 0, align)[1]
    buf = np.zeros((k, per))
    buf[:, :local.shape[1]] = local
    outs = [torch.zeros((k, per), dtype=torch.float64) for _ in range(world)]
    td.all_gather(outs, torch.from_numpy(buf))
    full = np.empty((k, n_vox))
    for r in range(world):
        c0, c1 = shard_columns(n_vox, world, r, align)
        full[:, c0:c1] = outs[r].numpy()[:, :c1 - c0]
    return full
