"""Rank bootstrap for the t-sharded multi-GPU solve: one process per GPU (launched e.g. by
`python -m torch.distributed.run --nproc-per-node N ...`, which only provides RANK / LOCAL_RANK /
WORLD_SIZE / MASTER_ADDR / MASTER_PORT), NCCL communicator created inside libflof_b200.so.

The 128-byte NCCL unique id is handed from rank 0 to the other ranks over a plain TCP socket on
MASTER_ADDR:(MASTER_PORT + 17).  torch is deliberately not imported here: the library links the
system libnccl.so.2, and loading torch's bundled NCCL into the same process first/second would mix
two NCCL versions behind one soname.
"""
import os
import socket
import time

from . import capi

ID_BYTES = 128
PORT_OFFSET = 17


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def exchange_unique_id(rank, world, make_id, timeout=120.0):
    addr = os.environ.get("MASTER_ADDR", "127.0.0.1")
    port = int(os.environ.get("MASTER_PORT", "29500")) + PORT_OFFSET
    if rank == 0:
        uid = make_id()
        srv = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        srv.bind(("0.0.0.0" if addr not in ("127.0.0.1", "localhost") else "127.0.0.1", port))
        srv.listen(world)
        srv.settimeout(timeout)
        conns = []
        for _ in range(world - 1):
            c, _a = srv.accept()
            conns.append(c)
        for c in conns:
            c.sendall(uid)
            c.close()
        srv.close()
        return uid
    t0 = time.time()
    while True:
        try:
            c = socket.create_connection((addr, port), timeout=5.0)
            break
        except OSError:
            if time.time() - t0 > timeout:
                raise
            time.sleep(0.2)
    buf = b""
    c.settimeout(timeout)
    while len(buf) < ID_BYTES:
        chunk = c.recv(ID_BYTES - len(buf))
        if not chunk:
            raise RuntimeError("rank 0 closed the rendezvous socket early")
        buf += chunk
    c.close()
    return buf


def init(ctx=None):
    """Creates (or completes) a Context for this rank and, if WORLD_SIZE > 1, its NCCL communicator."""
    import ctypes as C
    rank, world, local = env_rank()
    ctx = ctx or capi.Context(local)
    if world > 1:
        def make_id():
            buf = C.create_string_buffer(ID_BYTES)
            ctx._chk(ctx.lib.flof_comm_unique_id(buf))
            return buf.raw
        uid = exchange_unique_id(rank, world, make_id)
        ctx._chk(ctx.lib.flof_ctx_comm_init(ctx.h, world, rank, C.create_string_buffer(uid, ID_BYTES)))
    return ctx, rank, world


def barrier(ctx):
    ctx._chk(ctx.lib.flof_comm_barrier(ctx.h))


def max_over_ranks(ctx, value):
    import ctypes as C
    v = C.c_double(float(value))
    ctx._chk(ctx.lib.flof_comm_allreduce_max_host(ctx.h, C.byref(v)))
    return v.value
