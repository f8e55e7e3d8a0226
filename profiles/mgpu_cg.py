"""Distributed CG on the column-partitioned device matrix (SURVEY 8e "CG consumer", BASELINE config 5 shape: 3-D cube, degree 4,
assembly across the GPUs followed by a CG solve).  torchrun / nccl, one process per GPU.

  --check : degree-3 cube, N-rank assembly + DistributedCG against a single-rank assembly + gsb200_cg_host on rank 0.
  default : degree --degree, --nelem^3 elements per GPU slab: assembly time, CG iterations/s, final residual.
"""
import argparse, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gismo_b200 as g
from gismo_b200 import host, distributed as D


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true"); ap.add_argument("--degree", type=int, default=4); ap.add_argument("--nelem", type=int, default=96)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--cube", type=int, default=0, help="total elements per direction of ONE cube split over the ranks (strong); 0: nelem^3 per rank (weak)")
    ap.add_argument("--as-rank", type=str, default="", help="debug: 'r/w' = build and assemble the share of rank r of w in this single process (no CG)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.as_rank:
        r, w = (int(v) for v in a.as_rank.split("/"))
        prog = g.expr_compile("3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)")
        t0 = time.time(); pb = host.poisson_box_problem(3, a.degree, [a.cube] * 3, prog, rank=r, nranks=w); print("host problem", time.time() - t0, "s, dofs", pb.nfree, flush=True)
        A = g.DeviceAssembler(pb, device=0); print("created", flush=True)
        nnz = A.buildPattern(); print("pattern nnz", nnz, torch.cuda.mem_get_info(), flush=True)
        for _ in range(4):
            A.assemble()
        tm = A.timings(); print("assemble ms", tm.total_ms, "chunks", tm.nchunks, [tm.geometry_ms, list(tm.sweep_ms), tm.rhs_ms], torch.cuda.mem_get_info(), flush=True)
        return
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream().cuda_stream
    prog = g.expr_compile("3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)")

    def assemble(p, nel, r, w):
        pb = host.poisson_box_problem(3, p, nel, prog, rank=r, nranks=w)
        A = g.DeviceAssembler(pb, device=local, stream=stream)
        A.assemble()
        v = A.device_view()
        b = D.device_tensor(v.rhs, pb.nfree, torch.float64, local).clone()
        if w > 1:
            dist.all_reduce(b)
        return pb, A, b

    if a.check:
        pb, A, b = assemble(3, [14, 14, 14 * world], rank, world)
        x, it, res = D.DistributedCG(A, local).solve(b, max_iter=2000, tol=1e-12)
        ok = True
        if rank == 0:
            pb1, A1, b1 = assemble(3, [14, 14, 14 * world], 0, 1)
            x1, it1, res1 = A1.cg(b1.cpu().numpy(), max_iter=2000, tol=1e-12)
            err = np.abs(x.cpu().numpy() - x1).max() / np.abs(x1).max()
            db = float((b - b1).abs().max() / b1.abs().max())
            ok = err < 1e-8 and db < 1e-13
            print(f"CHECK distributed CG: {world} rank(s), {pb.nfree} DOFs, {it} iterations (single rank: {it1}), rel.res {res:.2e}, "
                  f"|x - x_single|/|x| = {err:.2e}, |b - b_single|/|b| = {db:.2e}: {'OK' if ok else 'FAIL'}", flush=True)
            A1.close()
        A.close()
    else:
        m = a.nelem
        t0 = time.time()
        shape = [a.cube] * 3 if a.cube else [m, m, m * world]
        pb = host.poisson_box_problem(3, a.degree, shape, prog, rank=rank, nranks=world)
        t_host = time.time() - t0
        A = g.DeviceAssembler(pb, device=local, stream=stream)
        nnz = A.buildPattern()
        for _ in range(3):
            A.assemble(sync=False)
        A.synchronize()
        if world > 1: dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); A.assemble(sync=False); e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        v = A.device_view()
        b = D.device_tensor(v.rhs, pb.nfree, torch.float64, local).clone()
        if world > 1:
            dist.all_reduce(b); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        cg = D.DistributedCG(A, local)
        torch.cuda.synchronize(); t1 = time.time()
        x, it, res = cg.solve(b, max_iter=a.iters, tol=1e-30)
        torch.cuda.synchronize(); t_cg = time.time() - t1
        if rank == 0:
            print(json.dumps({"config": f"3-D cube p={a.degree}, {shape[0]}x{shape[1]}x{shape[2]} elements", "host_problem_s": t_host, "chunks": A.timings().nchunks, "n_gpus": world, "dofs": pb.nfree, "nnz_per_gpu": nnz,
                              "assemble_ms": float(ms.item()), "assembled_dofs_per_sec": pb.nfree / (float(ms.item()) * 1e-3),
                              "cg_iterations": it, "cg_ms_per_iteration": t_cg / it * 1e3, "cg_rel_residual": res}), flush=True)
        A.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
