"""Multi-GPU bootstrap: one process per GPU (torchrun-style environment), NCCL underneath.

The data path (packed-halo ncclAllGather before a sharded SpMV, ncclAllReduce of the
fused inner products) lives in libkrylov_b200 (csrc/comm.cu).  This module only

  * reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_PORT from the environment,
  * gets the 128-byte NCCL unique id from rank 0 to everybody through a small
    file rendezvous on the node-local filesystem (single node),
  * splits rows into contiguous blocks (1-D row sharding).
"""
import os
import tempfile
import time

from . import _lib as L

__all__ = ["env_world", "row_partition", "exchange_unique_id", "init_from_env"]


def env_world():
    """(rank, world_size, local_rank) from the launcher's environment (defaults: single process)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def row_partition(n, nranks):
    """Contiguous row blocks: list of (row_begin, row_end), sizes differ by at most one."""
    base, extra = divmod(int(n), int(nranks))
    bounds, start = [], 0
    for r in range(nranks):
        size = base + (1 if r < extra else 0)
        bounds.append((start, start + size))
        start += size
    return bounds


def _rendezvous_path(tag=None):
    if tag is None:
        tag = os.environ.get("KRY_RENDEZVOUS_TAG")     # explicit (e.g. ranks started under different parents)
    if tag is None:
        # all workers of one torchrun share the agent as parent and the master port
        tag = "%s_%s_%d" % (os.environ.get("MASTER_PORT", "0"),
                            os.environ.get("TORCHELASTIC_RUN_ID", "none"), os.getppid())
    return os.path.join(tempfile.gettempdir(), "krylov_b200_ncclid_%s" % tag)


def exchange_unique_id(rank, nranks, make_id, tag=None, timeout=120.0):
    """Rank 0 calls ``make_id()`` and publishes the bytes; the others poll for them."""
    path = _rendezvous_path(tag)
    if rank == 0:
        uid = make_id()
        tmp = path + ".tmp.%d" % os.getpid()
        with open(tmp, "wb") as fh:
            fh.write(uid)
        os.replace(tmp, path)                      # atomic publish
        return uid
    deadline = time.time() + timeout
    while time.time() < deadline:
        try:
            with open(path, "rb") as fh:
                uid = fh.read()
            if len(uid) == L.KRY_COMM_ID_BYTES:
                return uid
        except OSError:
            pass
        time.sleep(0.01)
    raise TimeoutError("rank %d: no NCCL unique id at %s after %.0f s" % (rank, path, timeout))


def _make_nccl_id():
    import ctypes as C
    buf = C.create_string_buffer(L.KRY_COMM_ID_BYTES)
    L.call("kry_comm_unique_id", buf)
    return buf.raw


def init_from_env(context=None, tag=None):
    """Create (or take) the context on cuda:LOCAL_RANK and join the NCCL communicator
    described by the environment.  Returns (context, rank, world_size)."""
    from .device import Context
    rank, world, local = env_world()
    ctx = context or Context(local)
    if world > 1:
        uid = exchange_unique_id(rank, world, _make_nccl_id, tag=tag)
        ctx.comm_init(world, rank, uid)
        ctx.barrier()
        if rank == 0:
            try:
                os.remove(_rendezvous_path(tag))
            except OSError:
                pass
    return ctx, rank, world
