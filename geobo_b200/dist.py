"""Host-side plumbing for the multi-GPU path: one process per GPU (torchrun or any launcher that sets RANK / WORLD_SIZE /
MASTER_ADDR / MASTER_PORT), the voxel columns of ``Pt = A.K`` sharded contiguously over ranks.

The control plane is a few lines of plain TCP (``socket``; rank 0 is the hub): it distributes the NCCL unique id, and gives
barriers, max-over-ranks of a timing and a small host all-gather to processes that have no NCCL communicator (the CPU tests).
There is no PyTorch anywhere in the package.  Every data-path collective (all-reduce of the AkA partial sums, Cholesky panel
broadcasts, all-gather of the result shards) runs inside ``libgeobo_b200.so`` over NCCL / NVLink.
"""
import os
import pickle
import socket
import struct
import time

import numpy as np

_state = {"rank": 0, "world": 1, "initialized": False, "hub": None, "peers": None}
_MAGIC = b"geobo_b200-ctl1"
_PORT_OFFSETS = range(17, 17 + 24)          # MASTER_PORT itself belongs to the launcher's store (torchrun); we try MASTER_PORT + 17 ...


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


def shard_bounds(n_vox, world, weights=None, align=128):
    """Contiguous voxel-column shards [(c0, c1)] for all ranks.  ``weights`` (one non-negative number per equal-sized block of
    columns, e.g. per voxel row of the cube) balances the CUMULATIVE WEIGHT instead of the column count: with zero-digit culling
    the projection work of a voxel column depends on how many voxel rows lie inside the covariance's reach, so the ranks that own
    the cube's edge rows would otherwise finish early and wait in the AkA all-reduce.  Boundaries are multiples of ``align``; every
    rank owns at least ``align`` columns.  ``weights=None``: the uniform split of ``shard_columns``."""
    if weights is None or world == 1:
        return [shard_columns(n_vox, world, r, align) for r in range(world)]
    w = np.asarray(weights, dtype=float)
    if w.ndim != 1 or w.size < 1 or (w < 0).any() or not w.sum() > 0:
        raise ValueError("weights must be a non-empty vector of non-negative numbers")
    if n_vox < world * align:
        raise ValueError("cube with %d voxels is too small to shard over %d ranks at alignment %d" % (n_vox, world, align))
    # cumulative weight as a piecewise-linear function of the column index
    edges = np.linspace(0.0, float(n_vox), w.size + 1)
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for k in range(1, world):
        c = float(np.interp(k * cum[-1] / world, cum, edges))
        c = int(round(c / align)) * align
        c = max(c, cuts[-1] + align)                       # at least one aligned block per rank ...
        c = min(c, n_vox - (world - k) * align)            # ... also for the ranks still to come
        cuts.append(c)
    cuts.append(n_vox)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


# ------------------------------------------------------------------ TCP control plane (rank 0 = hub)
def _send(sock, obj):
    data = pickle.dumps(obj, protocol=4)
    sock.sendall(struct.pack("<Q", len(data)) + data)


def _recv_exact(sock, n):
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("geobo_b200.dist: control-plane peer closed the connection")
        buf += chunk
    return bytes(buf)


def _recv(sock):
    (n,) = struct.unpack("<Q", _recv_exact(sock, 8))
    return pickle.loads(_recv_exact(sock, n))


def _job_token(world):
    return "%s|%s|%d" % (os.environ.get("TORCHELASTIC_RUN_ID", ""), os.environ.get("MASTER_PORT", ""), world)


def _connect_plane(rk, world, timeout=180.0):
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    base = int(os.environ.get("MASTER_PORT", "29500"))
    token = _job_token(world)
    if rk == 0:
        srv = None
        for off in _PORT_OFFSETS:
            try:
                srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
                srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
                srv.bind((addr if addr not in ("localhost",) else "127.0.0.1", base + off))
                break
            except OSError:
                srv.close()
                srv = None
        if srv is None:
            raise RuntimeError("geobo_b200.dist: no free control-plane port next to MASTER_PORT=%d" % base)
        srv.listen(world + 8)
        srv.settimeout(timeout)
        peers = {}
        while len(peers) < world - 1:
            conn, _ = srv.accept()
            conn.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
            conn.settimeout(timeout)
            try:
                hello = _recv(conn)
            except Exception:
                conn.close()
                continue
            if not (isinstance(hello, tuple) and len(hello) == 3 and hello[0] == _MAGIC and hello[1] == token and 0 < hello[2] < world):
                conn.close()                      # a stranger, or a rank of another job probing the port range
                continue
            _send(conn, (_MAGIC, token))
            peers[hello[2]] = conn
        srv.close()
        _state["peers"] = [peers[r] for r in range(1, world)]
        for c in _state["peers"]:
            c.settimeout(None)
        return
    deadline = time.time() + timeout
    while time.time() < deadline:
        for off in _PORT_OFFSETS:
            try:
                s = socket.create_connection((addr, base + off), timeout=2.0)
            except OSError:
                continue
            try:
                s.setsockopt(socket.IPPROTO_TCP, socket.TCP_NODELAY, 1)
                s.settimeout(10.0)
                _send(s, (_MAGIC, token, rk))
                if _recv(s) == (_MAGIC, token):
                    s.settimeout(None)
                    _state["hub"] = s
                    return
            except Exception:
                pass
            s.close()
        time.sleep(0.05)
    raise RuntimeError("geobo_b200.dist: rank %d could not reach the control plane of rank 0 at %s:%d+" % (rk, addr, base + _PORT_OFFSETS[0]))


def _allgather_obj(obj):
    """Every rank contributes one picklable object; every rank receives the list ordered by rank (hub = rank 0)."""
    world, rk = _state["world"], _state["rank"]
    if world <= 1:
        return [obj]
    if rk == 0:
        objs = [obj] + [_recv(c) for c in _state["peers"]]
        for c in _state["peers"]:
            _send(c, objs)
        return objs
    _send(_state["hub"], obj)
    return _recv(_state["hub"])


def broadcast_bytes(data=None):
    """Rank 0's bytes on every rank."""
    return _allgather_obj(data if _state["rank"] == 0 else None)[0]


def init_from_env(ctx=None):
    """Join the job described by RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT (torchrun).  With
    WORLD_SIZE <= 1 this is a no-op.  If ``ctx`` (a ``_lib.Context``) is given, its NCCL communicator
    is created from a unique id broadcast by rank 0."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rk = int(os.environ.get("RANK", "0"))
    _state.update(rank=rk, world=world)
    if world <= 1:
        return rk, world
    if not _state["initialized"]:
        _connect_plane(rk, world)
        _state["initialized"] = True
    if ctx is not None and ctx.nranks != world:
        uid = broadcast_bytes(ctx.comm_unique_id() if rk == 0 else None)
        ctx.comm_init(uid, rk, world)
    return rk, world


def shutdown():
    for c in (_state["peers"] or []) + ([_state["hub"]] if _state["hub"] else []):
        try:
            c.close()
        except Exception:
            pass
    _state.update(rank=0, world=1, initialized=False, hub=None, peers=None)


def barrier():
    if _state["world"] > 1:
        _allgather_obj(None)


def max_over_ranks(value):
    if _state["world"] <= 1:
        return float(value)
    return float(max(_allgather_obj(float(value))))


def allgather_columns(local, n_vox, align=128, ctx=None, bounds=None):
    """``local``: (k, ncol_local) array of this rank's voxel columns -> (k, n_vox) on every rank.
    With a ``_lib.Context`` that has an NCCL communicator the gather runs over NCCL/NVLink inside the library
    (``gb_comm_allgather``); otherwise over the host control plane.  ``bounds``: the shards of all ranks (``shard_bounds``) when
    they are not the uniform split."""
    world, rk = _state["world"], _state["rank"]
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world <= 1:
        return local
    k = local.shape[0]
    if bounds is None:
        bounds = [shard_columns(n_vox, world, r, align) for r in range(world)]
    per = max(c1 - c0 for c0, c1 in bounds)
    if ctx is not None and getattr(ctx, "nranks", 1) == world:
        buf = np.zeros((k, per))
        buf[:, :local.shape[1]] = local
        outs = ctx.allgather(buf).reshape(world, k, per)
    else:
        outs = _allgather_obj(local)
    full = np.empty((k, n_vox))
    for r in range(world):
        c0, c1 = bounds[r]
        full[:, c0:c1] = outs[r][:, :c1 - c0]
    return full
