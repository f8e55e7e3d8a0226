"""Multi-GPU runs of the multi-patch configurations: one process per GPU (torchrun, nccl),
patches dealt round robin, interior columns owned outright, the coupled interface block and the rhs summed with ONE all_reduce
over NVLink on the library's own device buffers (gismo_b200/distributed.py).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/mgpu_configs.py [--check] [--config 3|4] ...

--check : parity of the N-rank result against the reference's fixtures (small problems), gathered on rank 0.
Timing  : CUDA events around assemble + exchange, max over ranks; prints one JSON line per configuration (rank 0).
"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import gismo_b200 as g
from gismo_b200 import capi, host, distributed as D


def exchange(A, pb, c0, outer_h, local):
    v = A.device_view()
    vals = D.device_tensor(v.values, int(v.nnz), torch.float64, local)
    rhs = D.device_tensor(v.rhs, pb.nfree * pb.nrhs, torch.float64, local)
    D.reduce_coupled_columns(vals, rhs, outer_h, c0)
    return vals, rhs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true"); ap.add_argument("--config", type=int, default=0)
    ap.add_argument("--nelem", type=int, default=0); ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream

    if a.check:
        import goldenutil as G
        for name in ["grid2x2_p2_m4", "grid2x2x2_p2_m3", "yeti_mp2_p2_m2", "elasticity_2cubes_p2"]:
            pb, z = G.load(name, g.expr_compile)
            pb.struct.rank, pb.struct.nranks = rank, world
            A = g.DeviceAssembler(pb, device=local, stream=stream)
            A.assemble()
            o, i, v = A.matrix()
            c0 = D.coupled_column_ranges(pb)
            vals, rhs = exchange(A, pb, c0, o, local)
            torch.cuda.synchronize()
            piece = (o, i, vals.cpu().numpy())
            pieces = [None] * world
            if world > 1:
                dist.all_gather_object(pieces, piece)
            else:
                pieces = [piece]
            if rank == 0:
                outer, inner, values = D.merge_rank_matrices(pieces, 0, pb.nfree)
                try:
                    G.check_against((outer, inner, values, rhs.cpu().numpy().reshape(pb.nrhs, pb.nfree).T), z, 1e-12)
                    print(f"CHECK {name}: {world} rank(s), nccl exchange of column runs {c0} OK", flush=True)
                except AssertionError as e:
                    print(f"CHECK {name}: FAIL {e}", flush=True)
            A.close()

    cfgs = []
    if a.config in (0, 3):   # config 3 shape: 2-D, degree 2, 21 glued patches (yeti_mp2 has 21), 512^2 elements each (r = 8)
        cfgs.append(("config3: 2-D multipatch p=2, 3x7 patches", dict(dim=2, degree=2, grid=(3, 7), nelem=a.nelem or 512, form=capi.FORM_POISSON,
                                                                       rhs=["2*pi^2*sin(pi*x)*sin(pi*y)"])))
    if a.config in (0, 4):   # config 4 shape: 3-D linear elasticity, degree 2, 2x2x2 patches of 75^3 elements, lambda = mu = 80000
        cfgs.append(("config4: 3-D elasticity p=2, 2x2x2 patches", dict(dim=3, degree=2, grid=(2, 2, 2), nelem=a.nelem or 75, form=capi.FORM_ELASTICITY,
                                                                         rhs=["0", "0", "-1000"], coef=(80000.0, 80000.0))))
    for title, c in cfgs:
        progs = [g.expr_compile(t) for t in c["rhs"]]
        t0 = time.time()
        pb = host.multipatch_grid_problem(c["dim"], c["degree"], c["grid"], c["nelem"], rhs_programs=progs, form=c["form"], coef=c.get("coef", (0.0, 0.0)),
                                          rank=rank, nranks=world)
        t_build = time.time() - t0
        A = g.DeviceAssembler(pb, device=local, stream=stream)
        t0 = time.time(); nnz = A.buildPattern(); torch.cuda.synchronize(); t_pat = time.time() - t0
        outer_h = np.zeros(pb.nfree + 1, np.int64)
        v = A.device_view()
        outer_h[:] = D.device_tensor(v.outer, pb.nfree + 1, torch.int64, local).cpu().numpy()
        c0 = D.coupled_column_ranges(pb)
        ncoupled = sum(b - a for a, b in c0)
        for _ in range(2):
            A.assemble(sync=False); exchange(A, pb, c0, outer_h, local)
        torch.cuda.synchronize()
        if world > 1: dist.barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        for _ in range(a.steps):
            A.assemble(sync=False)
        e1.record()
        for _ in range(a.steps):
            exchange(A, pb, c0, outer_h, local)
        e2.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.steps, e1.elapsed_time(e2) / a.steps], device="cuda")
        nnz_t = torch.tensor([float(nnz)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX); dist.all_reduce(nnz_t, op=dist.ReduceOp.MAX)
        A.synchronize()
        tm = A.timings()
        if rank == 0:
            nel = int(np.prod(c["grid"])) * c["nelem"] ** c["dim"]
            tot = float(ms[0] + ms[1])
            print(json.dumps({"config": title, "n_gpus": world, "nelem_per_patch_dir": c["nelem"], "dofs": pb.nfree, "elements": nel,
                              "nnz_max_per_rank": int(nnz_t.item()), "coupled_columns": ncoupled,
                              "assemble_ms": float(ms[0]), "exchange_ms": float(ms[1]), "dofs_per_sec": pb.nfree / (tot * 1e-3),
                              "qp_per_sec": nel * (c["degree"] + 1) ** c["dim"] / (tot * 1e-3), "pattern_s": t_pat, "host_problem_s": t_build,
                              "stages_rank0_ms": {"geometry": tm.geometry_ms, "sweeps": list(tm.sweep_ms), "rhs": tm.rhs_ms}}), flush=True)
        A.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
