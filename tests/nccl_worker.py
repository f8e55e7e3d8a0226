"""Worker of tests/test_gpu_parity.py::test_nccl_* (one process per GPU, started with torch.distributed.run): the
library's own NCCL path - gsb200_comm_init, gsb200_exchange (K4), gsb200_cg_solve - against the reference fixtures.
Prints one line 'NCCLWORKER ok ...' on rank 0; any failure exits non-zero."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")            # host transport only ships the NCCL id and gathers pieces for the check
    import goldenutil as G
    import gismo_b200 as g
    from gismo_b200 import distributed as D
    report = []
    for name in sys.argv[1:]:
        pb, z = G.load(name, g.expr_compile)
        pb = pb.with_fixed(pb.fixed, rank=rank, nranks=world)
        A = g.DeviceAssembler(pb, device=local)
        D.init_comm(A)                         # gsb200_comm_unique_id on rank 0 -> gsb200_comm_init everywhere
        A.assemble()
        A.exchange()
        nbytes, ncalls = A.comm_stats()
        o, i, v = A.matrix()
        b = A.rhs()
        x, it, res = A.cg_solve(max_iter=3000, tol=1e-11, check_every=10)
        cg_bytes, _ = A.comm_stats()
        view = A.device_view()
        pieces = [None] * world
        dist.all_gather_object(pieces, (o, i, v, view.col_begin, view.col_end))
        if rank == 0:
            outer, inner, values = D.merge_rank_matrices([p[:3] for p in pieces], pb.nfree)
            G.check_against((outer, inner, values, b), z, 1e-12)
            import scipy.sparse as sp
            K = sp.csc_matrix((values, inner, outer), shape=(pb.nfree, pb.nfree))
            r = K @ x - b[:, 0]
            rel = np.linalg.norm(r) / np.linalg.norm(b[:, 0])
            assert res <= 1e-11 and rel <= 1e-9, (name, it, res, rel)
            if len(pb.patches) == 1:           # slabs: contiguous, disjoint, covering
                ext = sorted((p[3], p[4]) for p in pieces)
                assert ext[0][0] == 0 and ext[-1][1] == pb.nfree and all(ext[k][1] == ext[k + 1][0] for k in range(world - 1)), ext
            report.append(f"{name}: exchange {nbytes} B, cg {it} it rel {rel:.1e} ({cg_bytes} B)")
        xs = [None] * world
        dist.all_gather_object(xs, x)
        assert all(np.array_equal(xs[0], xk) for xk in xs), "ranks ended with different solutions"
        A.close()
    if rank == 0:
        print("NCCLWORKER ok | " + " | ".join(report), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
