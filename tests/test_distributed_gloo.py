"""N>1 path on CPU: world_size-2 gloo process groups exercising the partition + exchange logic of
the library's exchange (gsb200_exchange, gsb200_cg_solve; the reduction goes through gsb200_set_allreduce -> gloo).  Per-rank assembly runs through the kernel interpreter (tests/emul, test
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
    import gismo_b200 as g
    from gismo_b200 import distributed as D
    pb, z = G.load(name, R.emul_compile)
    pb = pb.with_fixed(pb.fixed, rank=rank, nranks=world)
    A = g.DeviceAssembler(pb, lib=R.emul_lib())
    D.use_torch_allreduce(A)                 # gsb200_set_allreduce: the library's K4 exchange over gloo
    A.assemble()
    A.exchange()                             # coupled columns + right-hand side
    o, i, v = A.matrix()
    b = A.rhs()
    nbytes, ncalls = A.comm_stats()
    # the CG consumer across the ranks (patch-wise ownership or slabs over the callback: full-length reduction of the product)
    x, it, res = A.cg_solve(max_iter=400, tol=1e-10, check_every=5)
    pieces = [None] * world
    dist.all_gather_object(pieces, (o, i, v))
    if rank == 0:
        outer, inner, values = D.merge_rank_matrices(pieces, pb.nfree)
        try:
            G.check_against((outer, inner, values, b), z, 1e-12)
            assert ncalls == 1 and nbytes >= 8 * pb.nfree
            import scipy.sparse as sp
            K = sp.csc_matrix((values, inner, outer), shape=(pb.nfree, pb.nfree))
            r = K @ x - b[:, 0]
            assert res <= 1e-10 and np.linalg.norm(r) <= 1e-8 * np.linalg.norm(b[:, 0]), (it, res, np.linalg.norm(r))
            out.put("ok")
        except AssertionError as e:
            out.put("FAIL " + str(e))
    A.close()
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
