"""N>1 path on CPU: world_size-2 gloo process groups exercising the partition + exchange logic of
gismo_b200/distributed.py.  Per-rank assembly runs through the kernel interpreter (tests/emul, test
harness) because this container has no GPU; on the B200 box the same code path runs with nccl
(bench.py --gpus N, tests/test_gpu_parity.py::test_rank_slabs_cover_the_matrix)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, name, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import goldenutil as G
    import refutil as R
    from gismo_b200 import distributed as D
    pb, z = G.load(name, R.emul_compile)
    pb.struct.rank, pb.struct.nranks = rank, world
    o, i, v, b, _ = R.lib_assemble(R.emul_lib(), pb)
    c0 = D.first_coupled_column(pb)
    vt, bt = torch.from_numpy(v), torch.from_numpy(np.ascontiguousarray(b[:, 0]))
    D.reduce_coupled_columns(vt, bt, o, c0)
    # gather every rank's piece on rank 0 and verify against the reference fixture
    pieces = [None] * world
    dist.all_gather_object(pieces, (o, i, vt.numpy()))
    if rank == 0:
        outer, inner, values = D.merge_rank_matrices(pieces, c0, pb.nfree)
        try:
            G.check_against((outer, inner, values, bt.numpy()[:, None]), z, 1e-12)
            out.put("ok")
        except AssertionError as e:
            out.put("FAIL " + str(e))
    dist.destroy_process_group()


@pytest.mark.parametrize("name", ["grid2x2_p2_m4", "grid2x2x2_p2_m3", "yeti_mp2_p2_m2", "cube_p3_curved_m4"])
def test_two_rank_partition_and_exchange(name):
    import refutil as R
    R.emul_lib()   # build once before forking
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, out)) for r in range(2)]
    [p.start() for p in procs]
    [p.join(120) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert out.get(timeout=5) == "ok"
