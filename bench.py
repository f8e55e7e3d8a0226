#!/usr/bin/env python
"""bench.py — assembled DOFs/s (and quadrature points/s) of isogeometric system assembly on B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W                 # config 2 (headline): 3-D p=3 Poisson, 125^3 elements per GPU
    python bench.py --config 3|4|5|target --gpus N ...            # the other BASELINE configs (see WORKLOADS below)
    python bench.py --impl reference --steps K --warmup W         # the reference's CPU assembler (OpenMP, all host cores)

One JSON line on stdout (rank 0).  `value` = whole-job DOFs/s with inputs resident in HBM (pattern built, tables uploaded;
the timed region is K calls of gsb200_assemble [+ gsb200_exchange when patches are sharded] on the device, CUDA-event timed, max
over ranks).  `e2e` = the same metric through the host-buffer entry points a gismo caller uses for a repeated assembly
(gsPoissonAssemblerB200::setKeepPattern): eliminated-DOF values up from pinned host memory, values + right-hand side down into
pinned host memory; the first assembly (problem upload, pattern, index arrays) is reported beside it.
N>1: one process per GPU.  Single patches are cut into slabs of matrix columns along the last direction (no data-path collective);
multi-patch configs shard whole patches and sum the coupled interface columns + rhs with the library's NCCL exchange (K4), which
is inside the timed region.  Every run verifies what it timed with an oracle-independent invariant (constant null space, see
invariant_check) and, for `--config 5|target`, solves the system with the device CG.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

HOST_CORES = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
if "reference" in sys.argv or "--cpu-matrix" in sys.argv:
    # the CPU baseline runs with bound threads.  Only in the reference's own process (the OpenMP runtime reads this
    # when it is loaded and pins the initial thread; the GPU arm's host threads must stay free), and with all host cores whatever
    # torchrun put into OMP_NUM_THREADS.
    os.environ.setdefault("OMP_PROC_BIND", "close")
    os.environ["OMP_NUM_THREADS"] = str(HOST_CORES)

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

F3 = "3*pi^2*sin(pi*x)*sin(pi*y)*sin(pi*z)"

WORKLOADS = {
    "2": "3D unit-cube tensor B-spline, degree 3, 125^3 elements per GPU slab (2.0 M DOFs per GPU), Poisson stiffness + RHS",
    "3": "2D multipatch domain filedata/domain2d/yeti_mp2.xml (21 patches, topology from the reference fixture), degree 2, 512^2 elements per patch (6.6 M DOFs), patches sharded, interface exchange",
    "4": "3D linear elasticity, degree 2, 2x2x2 patches of 75^3 elements (10.3 M DOFs), patches sharded, interface exchange",
    "5": "3D unit-cube, degree 4, 368x368x(46 N) elements (51.5 M DOFs at 8 GPUs), Poisson assembly + CG solve",
    "target": "3D unit-cube, degree 3, 368x368x(46 N) elements (51.0 M DOFs at 8 GPUs), Poisson assembly + CG solve",
}


def f_sf(p, dim):
    """sum-factorised element-wise flop count per element (no symmetry)."""
    q = p + 1
    return q ** 3 * (18 * q ** 4 + 24 * q ** 3) if dim == 3 else q ** 2 * (8 * q ** 3 + 12 * q ** 2)


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                              timeout=5).decode().strip()
                self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def host_cores():
    return HOST_CORES


def cpu_reference_matrix(degree, nelem_all, nelem_one):
    """both reference paths (gsPoissonAssembler = visitor, gsExprAssembler = expression) x {1, all} threads,
    OMP_PROC_BIND=close, on bounded samples of the config-2 workload.  DOFs/s each."""
    import refutil as R
    cores = host_cores()
    out = {}
    for path, pname in ((0, "gsPoissonAssembler"), (1, "gsExprAssembler")):
        for thr, m in ((cores, nelem_all), (1, nelem_one)):
            ref = R.ref_run(dim=3, degree=degree, nelem=m, geometry=0, path=path, rhs=[F3], dir_values=100, threads=thr)
            out[f"{pname}_{'all' if thr == cores else '1'}"] = {"value": ref.nfree / ref.seconds, "cores": thr, "seconds": ref.seconds,
                                                              "sample": f"3D p={degree}, {m}^3 elements, {ref.nfree} DOFs"}
    return out


def cpu_matrix_arm(args):
    """bench.py --impl reference --cpu-matrix: prints the four-way CPU matrix as JSON (run as a subprocess of the GPU arm)."""
    import refutil as R
    if not R.have_ref():
        print(json.dumps({"unavailable": "oracle/_ref/libgsref.so did not travel"}), flush=True)
        return
    print(json.dumps(cpu_reference_matrix(args.degree or 3, args.ref_nelem, args.ref_nelem)), flush=True)       # the same sample for every entry


def reference_arm(args, rank):
    """The reference's own CPU assembler (gsPoissonAssembler::assemble, OpenMP over all host cores whatever torchrun put into
    OMP_NUM_THREADS) on a bounded sample of the workload."""
    if rank != 0:
        return
    import refutil as R
    m, deg = args.ref_nelem, args.degree or 3
    times = []
    if R.have_ref():
        kind, cores = "reference", host_cores()
        for it in range(args.warmup + args.steps):
            ref = R.ref_run(dim=3, degree=deg, nelem=m, geometry=0, rhs=[F3], dir_values=100, threads=cores)
            if it >= args.warmup:
                times.append(ref.seconds)
        ndof, nqp = ref.nfree, ref.qpoints
    else:  # reference build did not travel: the C restatement (single thread)
        import gismo_b200 as g
        kind, cores = "port", 1
        pb = g.host.poisson_box_problem(3, deg, m, R.emul_compile(F3))
        for it in range(args.warmup + args.steps):
            t0 = time.time(); R.oracle_assemble(pb); dt = time.time() - t0
            if it >= args.warmup:
                times.append(dt)
        ndof, nqp = pb.nfree, m ** 3 * (deg + 1) ** 3
    t = float(np.mean(times))
    val = ndof / t
    sample = f"3D p={deg} unit cube, {m}^3 elements, {ndof} DOFs per step (bounded sample of the config-2 workload)"
    line = {"impl": "reference", "metric": "assembled_dofs_per_sec", "value": val, "unit": "DOFs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "qp_per_sec": nqp / t,
            "config": {"workload": f"3D unit-cube tensor B-spline, degree {deg}, Poisson stiffness + RHS (gsPoissonAssembler::assemble, OpenMP, OMP_PROC_BIND=close)", "sample": sample},
            "cpu_baseline": {"value": val, "unit": "DOFs/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": val, "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def build_workload(cfg, args, rank, world, g):
    """-> dict(pb, dim, p, n_elem (global), scaling, sharded (patches over ranks), desc)."""
    import goldenutil as G
    host = g.host
    if cfg == "2":
        p, m = args.degree or 3, args.nelem or 125
        pb = host.poisson_box_problem(3, p, [m, m, m * world], g.expr_compile(F3), rank=rank, nranks=world)
        return dict(pb=pb, dim=3, p=p, n_elem=m * m * m * world, scaling="weak", sharded=False,
                    desc=f"3D unit-cube tensor B-spline, degree {p}, {m}x{m}x{m * world} elements ({m}^3 per GPU slab)")
    if cfg in ("5", "target"):
        p = args.degree or (4 if cfg == "5" else 3)
        m, mz = args.nelem or 368, (args.nelem_z or 46) * world
        pb = host.poisson_box_problem(3, p, [m, m, mz], g.expr_compile(F3), rank=rank, nranks=world)
        return dict(pb=pb, dim=3, p=p, n_elem=m * m * mz, scaling="weak", sharded=False,
                    desc=f"3D unit-cube tensor B-spline, degree {p}, {m}x{m}x{mz} elements ({m}x{m}x{mz // world} per GPU slab)")
    if cfg == "3":
        p, ne = args.degree or 2, args.nelem or 256
        small, _ = G.load("yeti_mp2_p2_m2", g.expr_compile)
        pb = host.refine_multipatch_2d(small, p, ne, rhs_programs=[g.expr_compile("1")], rank=rank, nranks=world)
        return dict(pb=pb, dim=2, p=p, n_elem=21 * (2 * ne) ** 2, scaling="strong", sharded=True,
                    desc=f"2D multipatch yeti_mp2.xml (21 patches, glued interfaces), degree {p}, {2 * ne}^2 elements per patch")
    if cfg == "4":
        p, ne = args.degree or 2, args.nelem or 75
        progs = [g.expr_compile(t) for t in ("x", "y*z", "1")]
        pb = host.multipatch_grid_problem(3, p, [2, 2, 2], ne, rhs_programs=progs, form=g.capi.FORM_ELASTICITY, coef=(2.0, 1.5),
                                          rank=rank, nranks=world)
        return dict(pb=pb, dim=3, p=p, n_elem=8 * ne ** 3, scaling="strong", sharded=True,
                    desc=f"3D linear elasticity (lambda 2, mu 1.5), degree {p}, 2x2x2 patches of {ne}^3 elements")
    raise SystemExit(f"unknown --config {cfg}")


def invariant_check(A, pb, torch, dist, world, local, exchange):
    """Oracle-independent check of EVERY matrix row at any size, on the device: B-splines are a partition of unity, so a constant
    field has zero gradient / zero strain: K_ff 1 + K_fe 1 = 0, i.e. the assembled (free x free) matrix times the constant
    vector must equal rhs(g = 1) - rhs(g = 0), the elimination terms of Dirichlet data 1 (each vector component in turn for
    elasticity).  Uses only product entry points: gsb200_set_fixed, gsb200_assemble (+ gsb200_exchange), gsb200_spmv_device.
    Also the symmetry residual x.(K y) - y.(K x) for two probe vectors.  Returns the largest relative deviations."""
    from gismo_b200 import distributed as D
    n, ncomp = pb.nfree, pb.ncomp
    v = A.device_view()
    dev = torch.device("cuda", local)

    def rhs_with(fixed):
        A.set_fixed(fixed)
        A.assemble(sync=False)
        if exchange:
            A.exchange()
        A.synchronize()
        return D.device_tensor(v.rhs, n, torch.float64, local).clone()

    def spmv(x):
        y = torch.zeros(n, dtype=torch.float64, device=dev)
        torch.cuda.synchronize()
        A._check(A.lib.gsb200_spmv_device(A._h, x.data_ptr(), y.data_ptr()))
        A.synchronize()
        return y
    outer = D.device_tensor(v.outer, n + 1, torch.int64, local)
    stored = (outer[1:] - outer[:-1]) > 0
    if world > 1 and not exchange:          # slabs: a rank's rhs holds its own rows only, which is all that is compared
        pass
    r0 = rhs_with(np.zeros((pb.nfixed, 1)))
    worst = 0.0
    per, nf1 = pb.nfixed // ncomp, n // ncomp                              # free and eliminated DOFs are numbered component-major
    for c in range(ncomp):
        fx = np.zeros((pb.nfixed, 1))
        fx[c * per:(c + 1) * per] = 1.0
        r1 = rhs_with(fx)
        e = torch.zeros(n, dtype=torch.float64, device=dev)
        e[c * nf1:(c + 1) * nf1] = 1.0
        y = spmv(e)
        if bool(stored.any()):
            scale = float(torch.max(torch.abs(D.device_tensor(v.values, int(v.nnz), torch.float64, local)[:: max(1, int(v.nnz) // 1000003)])))
            worst = max(worst, float(torch.max(torch.abs((y - (r1 - r0))[stored]))) / max(scale, 1e-300))
    A.set_fixed(np.zeros((pb.nfixed, 1)))
    idx = torch.arange(n, dtype=torch.float64, device=dev)
    xa, xb = torch.cos(0.37 * idx + 0.11), torch.sin(0.11 * idx + 0.5)
    ya, yb = spmv(xa), spmv(xb)
    if exchange:        # coupled columns are stored by every rank: count them once
        hold = stored.to(torch.float64)
        dist.all_reduce(hold)
        wgt = torch.where(stored, 1.0 / torch.clamp(hold, min=1.0), torch.zeros_like(hold))
    else:
        wgt = stored.to(torch.float64)
    s = torch.stack([torch.dot(xb * wgt, ya), torch.dot(xa * wgt, yb), torch.dot(ya * wgt, ya), torch.dot(yb * wgt, yb)])
    if world > 1:
        dist.all_reduce(s)
    sym = abs(float(s[0] - s[1])) / max((float(s[2]) * float(s[3])) ** 0.5, 1e-300)
    w = torch.tensor([worst, sym], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(w, op=dist.ReduceOp.MAX)
    return {"constant_nullspace_max_rel": float(w[0]), "symmetry_rel": float(w[1]), "rows_checked": "all stored rows of every rank",
            "bar": 1e-12, "ok": bool(float(w[0]) <= 1e-12 and float(w[1]) <= 1e-12)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="2", choices=sorted(WORKLOADS))
    ap.add_argument("--degree", type=int, default=0)
    ap.add_argument("--nelem", type=int, default=0, help="elements per direction (per GPU slab / per patch), 0 = the config's size")
    ap.add_argument("--nelem-z", type=int, default=0, help="configs 5/target: element layers per GPU along the last direction")
    ap.add_argument("--ref-nelem", type=int, default=24, help="elements per direction of the CPU sample")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cg-iters", type=int, default=4000)
    ap.add_argument("--cg-tol", type=float, default=1e-8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-matrix", action="store_true", help="with --impl reference: both reference paths x {1, all} threads as one JSON object")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="config 2: skip the short runs of configs 3 and 4 appended under `configs`")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if args.cpu_matrix:
            cpu_matrix_arm(args)
        else:
            reference_arm(args, rank)
        return

    # exactly ONE line on stdout: libraries print banners there (NCCL: "NCCL version ..."), so everything goes to stderr until the
    # JSON line is written
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    import gismo_b200 as g
    from gismo_b200 import distributed as D
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    stream = torch.cuda.current_stream().cuda_stream
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")

    def run_config(cfg, steps, warmup, full):
        """Device-resident timing (+ roofline when `full`) of one workload; returns the fields of its JSON object."""
        t_build = time.perf_counter()
        W = build_workload(cfg, args, rank, world, g)
        pb, p, dim = W["pb"], W["p"], W["dim"]
        t_build = time.perf_counter() - t_build
        exchange = W["sharded"] and world > 1
        t0 = time.perf_counter()
        A = g.DeviceAssembler(pb, device=local, stream=stream)
        t_create = time.perf_counter() - t0
        if world > 1:
            D.init_comm(A)                                     # the library's own NCCL communicator (gsb200_comm_init)
        t0 = time.perf_counter()
        nnz_local = A.buildPattern()
        t_pattern = time.perf_counter() - t0

        def step():
            A.assemble(sync=False)
            if exchange:
                A.exchange()
        for _ in range(max(warmup, 3)):
            step()
        A.synchronize()
        sampler = ClockSampler(local) if full else None
        if sampler:
            sampler.start()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        barrier()
        A.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        if sampler:
            sampler.stop_flag = True
        tm = A.timings()
        jit_used = A.jit_launches()
        xbytes, _ = A.comm_stats() if exchange else (0, 0)
        n_dofs, n_elem = pb.nfree, W["n_elem"]
        qp = n_elem * (p + 1) ** dim
        sec = float(ms.item()) / 1e3 / steps
        res = {"metric": "assembled_dofs_per_sec", "value": n_dofs / sec, "unit": "DOFs/s", "n_gpus": world, "steps": steps,
               "warmup": max(warmup, 3), "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": W["scaling"],
               "vs_baseline": None, "dtype": "f64", "data": "synthetic", "qp_per_sec": qp / sec,
               "config": {"workload": f"{W['desc']}, {n_dofs} DOFs, nnz on rank 0 {nnz_local}, {'elasticity' if pb.form == 1 else 'Poisson'} stiffness + RHS",
                          "baseline_config": cfg,
                          "l2": "no flush needed: every step streams tens of GB of sum-factorisation intermediates through HBM per GPU, far more than the 126 MB L2",
                          "parallelism": (f"{len(pb.patches)} patches over {world} ranks (longest first), coupled interface columns + rhs summed by gsb200_exchange (NCCL, in the timed region)"
                                          if W["sharded"] else f"column slabs x{world}, no collective in the data path"),
                          "source_term": ("compiled into the fused geometry + first-sweep kernel with NVRTC (repeated assemblies, from the 3rd use on; bitwise the "
                                          "interpreter's operations)" if jit_used else "interpreted stack machine")},
               "nccl_bytes_per_step": int(xbytes), "gpu_launches": int(tm.launches) * steps,
               "setup_ms": {"host_problem_build": t_build * 1e3, "create": t_create * 1e3, "pattern": t_pattern * 1e3}}
        stages = {"geometry_ms": tm.geometry_ms, "sweep_ms": [tm.sweep_ms[k] for k in range(3)], "rhs_ms": tm.rhs_ms,
                  "total_ms_last_step": tm.total_ms, "pattern_ms": tm.pattern_ms, "chunks": tm.nchunks,
                  "sweep_gbs": [tm.sweep_bytes[k] / (tm.sweep_ms[k] * 1e-3) / 1e9 if tm.sweep_ms[k] > 0 else 0 for k in range(3)],
                  "sweep_tflops": [tm.sweep_flops[k] / (tm.sweep_ms[k] * 1e-3) / 1e12 if tm.sweep_ms[k] > 0 else 0 for k in range(3)],
                  "executed_flops": int(sum(tm.sweep_flops))}
        # ---------------- verification of what was timed
        res["verification"] = invariant_check(A, pb, torch, dist, world, local, exchange)
        if cfg == "3" and (args.nelem or 256) == 8 and world == 1:      # the reference's own fingerprints exist at this size
            import goldenutil as G
            _, z = G.load("yeti_mp2_p2_m8", g.expr_compile)
            A.assemble()
            G.check_against(A.matrix() + (A.rhs(),), z, 1e-12)
            res["verification"]["reference_fixture"] = "yeti_mp2_p2_m8: pattern bit-exact, fingerprints within 1e-12"
        # ---------------- CG consumer (configs 5 / target: "followed by CG solve for validation")
        if cfg in ("5", "target") or (full and cfg == "2" and world == 1):
            A.set_fixed(np.zeros((pb.nfixed, 1)))
            step(); A.synchronize()
            if world > 1 and not exchange:
                A.exchange()                                   # slabs: every rank needs the whole right-hand side once
            iters_cap = args.cg_iters if cfg != "2" else 400
            barrier(); t0 = time.perf_counter()
            _, it, relres = A.cg_solve(None, max_iter=iters_cap, tol=args.cg_tol, check_every=25, want_x=False)
            barrier(); dt = time.perf_counter() - t0
            cbytes, _ = A.comm_stats()
            nreg, ntab = A.spmv_info()
            loop_ms, _ = A.cg_info()
            res["cg"] = {"iterations": it, "rel_residual": relres, "tol": args.cg_tol, "converged": bool(relres <= args.cg_tol), "seconds_with_setup": dt,
                         "ms_per_iteration": loop_ms / max(it, 1),
                         "matrix_GBps_per_gpu": 8.0 * nnz_local / (loop_ms * 1e-3 / max(it, 1)) / 1e9,
                         "timing": "device time of the iteration loop (CUDA events inside gsb200_cg_solve); seconds_with_setup adds work vectors, NCCL's lazy connections and the gathering of the solution",
                         "regular_columns_rank0": int(nreg), "nccl_bytes": int(cbytes),
                         "solver": "gsb200_cg_solve: Jacobi-preconditioned CG, device scalars, " + ("neighbour halo exchange of the search direction" if world > 1 and not W["sharded"] else "single rank" if world == 1 else "full-length reduction of the product")}
        if not full:
            A.close()
            return res, None
        # ---------------- roofline of the dominant kernel (longest sweep), measured in this run
        sweeps = [(tm.sweep_ms[k], k) for k in range(3)]
        dom_ms, dom = max(sweeps)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"sweep{dom}")
        except Exception:
            pass
        ach = tm.sweep_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        lanes = dom_ms <= 0          # multi-patch: the patches' kernels overlap on several streams, there are no per-sweep times
        if lanes:
            ach = sum(tm.sweep_bytes) / sec / 1e9
        kname = {0: "k_geo_sweep (geometry + source term + sweep of direction 0 fused; D and F stay in shared memory)" if dim == 3 else "first sweep",
                 1: "k_sweepw, sum-factorisation sweep of direction 1", 2: "k_sweepw, final sweep of direction 2 (CSC scatter)"}[dom]
        if lanes:
            kname, traffic = "all sweeps of the step (patches run concurrently on lanes: whole-step figure)", None
        roofline = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s",
                    "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(sum(tm.sweep_bytes) if lanes else tm.sweep_bytes[dom]),
                    "ms_per_launch": (sec * 1e3 if lanes else dom_ms / max(tm.nchunks, 1)),
                    "note": "achieved = algorithmic bytes of the sweep (its inputs read once + outputs written once, DESIGN.md 3) / its "
                            "event-timed duration; traffic = measured dram read+write bytes of the same launches (profiles/traffic.json). "
                            "per_sweep lists all three; step_vs_* put the WHOLE step against the executed flops and against the compulsory "
                            "output bytes of SURVEY 8(d), so that the sweep fraction cannot be misread as the step's",
                    "per_sweep": [{"ms": tm.sweep_ms[k], "GBps": stages["sweep_gbs"][k], "frac": stages["sweep_gbs"][k] / hbm_peak} for k in range(3)]}
        n_local_cols = int(round(pb.nfree / world))
        comp_bytes = 8 * nnz_local + 8 * n_local_cols
        if rank == 0:
            try:
                pk = g.measure_peaks(local)
                stages["measured_fp64_tflops"] = pk["fp64_tflops"]; stages["measured_dmma_tflops"] = pk["dmma_tflops"]; stages["measured_copy_gbs"] = pk["hbm_gbs"]
                fsf = f_sf(p, dim) * (n_elem / world)
                t_roof = max(fsf / (pk["fp64_tflops"] * 1e12), comp_bytes / (hbm_peak * 1e9))
                t_exec = max(stages["executed_flops"] / (pk["fp64_tflops"] * 1e12), comp_bytes / (hbm_peak * 1e9))
                roofline["step_vs_executed_flops_and_compulsory_bytes"] = {
                    "bound_ms": t_exec * 1e3, "achieved_ms": sec * 1e3, "frac": t_exec / sec,
                    "note": "max(executed flops / measured FP64 peak, (8 nnz + rhs) / HBM peak): the bound of THIS algorithm"}
                roofline["step_vs_compulsory_bytes"] = {"bytes": comp_bytes, "bound_ms": comp_bytes / (hbm_peak * 1e9) * 1e3, "frac": comp_bytes / (hbm_peak * 1e9) / sec}
                stages["assembly_roofline"] = {"F_SF_flops_per_gpu": fsf, "compulsory_bytes_per_gpu": comp_bytes, "roofline_ms": t_roof * 1e3,
                                               "achieved_ms": sec * 1e3, "frac": t_roof / sec,
                                               "note": "SURVEY 8(d) yardstick; global sum factorisation executes fewer flops than the element-wise F_SF count, so frac may exceed 1"}
            except Exception as e:  # noqa
                stages["peaks_error"] = str(e)
        res["roofline"] = roofline
        res["stages"] = stages
        res["clocks"] = sampler.summary()
        A.close()
        return res, (pb, nnz_local)

    # ================= main measurement
    res, keep = run_config(args.config, args.steps, args.warmup, full=True)
    pb, nnz_local = keep
    n_dofs = pb.nfree

    # ---------------- end to end through host buffers (pinned)
    e2e = None
    single = len(pb.patches) == 1
    if not args.no_e2e and nnz_local < 2 ** 31:
        outer = torch.empty(pb.nfree + 1, dtype=torch.int32).pin_memory().numpy()
        inner = torch.empty(max(nnz_local, 1), dtype=torch.int32).pin_memory().numpy()
        values = torch.empty(max(nnz_local, 1), dtype=torch.float64).pin_memory().numpy()
        rhs_h = torch.empty(pb.nfree * pb.nrhs, dtype=torch.float64).pin_memory().numpy()
        h2d_first = sum(pa.dofmap.nbytes + pa.geo_coefs.nbytes + sum(k.nbytes for k in pa.space_knots) + sum(k.nbytes for k in pa.geo_knots)
                        for pa in pb.patches) + sum(pr.ops.nbytes + pr.consts.nbytes for pr in pb.rhs_programs)
        d2h_first = 8 * (pb.nfree + 1) + 4 * nnz_local + 8 * nnz_local + 8 * pb.nfree
        first_t, first_phases = [], []
        B = None
        for it in range(2):
            if B is not None:
                B.close()
            barrier()
            t0 = time.perf_counter()
            B = g.DeviceAssembler(pb, device=local, stream=stream)      # H2D of the flattened problem, 1-D tables
            t1 = time.perf_counter()
            B.buildPattern()                                            # sparsity pattern on the device
            t2 = time.perf_counter()
            B.assemble_into(outer, inner, values, rhs_h)                # assembly + D2H of the CSC triple and rhs (pipelined)
            barrier()
            t3 = time.perf_counter()
            first_t.append(t3 - t0)
            first_phases.append([t1 - t0, t2 - t1, t3 - t2])
        sharded = world > 1 and not single
        if sharded:
            D.init_comm(B)
        # the repeated step of a gismo caller (gsPoissonAssemblerB200::setKeepPattern): new eliminated-DOF values go up from pinned
        # memory, values + right-hand side come back into pinned memory; the index arrays were delivered by the first assembly
        fixed_h = torch.zeros(max(pb.nfixed, 1), dtype=torch.float64).pin_memory().numpy()
        e2e_t = []
        for it in range(1 + args.e2e_steps):
            barrier()
            t0 = time.perf_counter()
            B.set_fixed(fixed_h[:pb.nfixed].reshape(pb.nfixed, 1))
            if sharded:                                                 # the coupled columns are exchanged on the device before they travel
                B.assemble(sync=False); B.exchange()
                B.matrix_into(outer, inner, values); B.rhs_into(rhs_h)
            else:
                B.assemble_values_into(values, rhs_h)
            barrier()
            if it > 0:
                e2e_t.append(time.perf_counter() - t0)
        tm_e2e = B.timings()
        # N > 1: the matrix stays where its consumer is (the device CG, gsb200_cg_solve): the hand-off step uploads the eliminated
        # values, assembles (+ exchanges) and reads back the right-hand side rows of the rank's own columns
        hand_t = []
        if world > 1:
            vw = B.device_view()
            c0, c1 = (vw.col_begin, vw.col_end) if single else (0, pb.nfree)
            rhs_dev = D.device_tensor(vw.rhs, pb.nfree, torch.float64, local)
            for it in range(1 + max(args.e2e_steps, 3)):
                barrier()
                t0 = time.perf_counter()
                B.set_fixed(fixed_h[:pb.nfixed].reshape(pb.nfixed, 1))
                B.assemble(sync=False)
                if sharded:
                    B.exchange()
                B.synchronize()
                torch.from_numpy(rhs_h)[c0:c1].copy_(rhs_dev[c0:c1], non_blocking=True)
                barrier()
                if it > 0:
                    hand_t.append(time.perf_counter() - t0)
        B.close()
        e2e_s = torch.tensor([float(np.mean(e2e_t)), float(first_t[-1]), float(np.mean(hand_t)) if hand_t else 0.0], device="cuda")
        if world > 1:
            dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        host_matrix = {"value": n_dofs / float(e2e_s[0].item()), "ms_per_step": float(e2e_s[0].item()) * 1e3,
                       "d2h_bytes_per_step": int((12 if sharded else 8) * nnz_local + 8 * pb.nfree)}
        if world > 1:
            e2e = {"value": n_dofs / float(e2e_s[2].item()), "unit": "DOFs/s", "h2d_bytes_per_step": int(8 * pb.nfixed), "d2h_bytes_per_step": int(8 * (c1 - c0)),
                   "steps": max(args.e2e_steps, 3), "ms_per_step": float(e2e_s[2].item()) * 1e3,
                   "includes": "N > 1: device hand-off to the device consumer (gsb200_cg_solve): gsb200_set_fixed from pinned host memory, gsb200_assemble "
                               "(+ gsb200_exchange), right-hand-side rows of the rank's columns into pinned host memory; the matrix values stay in HBM. "
                               "`matrix_to_host` is the N = 1 definition (every rank ships its values to the host): on this box the ranks share the "
                               "host's PCIe / memory path, so that number does not scale",
                   "matrix_to_host": host_matrix}
        else:
            e2e = None
        e2e1 = {"value": n_dofs / float(e2e_s[0].item()), "unit": "DOFs/s", "h2d_bytes_per_step": int(8 * pb.nfixed),
               "d2h_bytes_per_step": int((12 if sharded else 8) * nnz_local + 8 * pb.nfree),
               "steps": args.e2e_steps, "ms_per_step": float(e2e_s[0].item()) * 1e3, "delivery_chunks": int(tm_e2e.nchunks),
               "d2h_GBps": ((12 if sharded else 8) * nnz_local + 8 * pb.nfree) / float(e2e_s[0].item()) / 1e9,      # the step is the PCIe transfer: this is the box's rate
               "includes": ("repeated assembly on a kept handle (gsb200_set_fixed + gsb200_assemble_values_to_host): eliminated-DOF values "
                            "from pinned host memory, all kernels, values + right-hand side into pinned host memory (finished column "
                            "ranges travel while later chunks integrate); the index arrays travelled with the first assembly; every rank "
                            "delivers its own columns to its own host buffers") if not sharded else
                           "gsb200_set_fixed, gsb200_assemble, gsb200_exchange (NCCL), download of the rank's CSC arrays + rhs into pinned host memory",
               "first_assembly": {"ms": float(e2e_s[1].item()) * 1e3, "value": n_dofs / float(e2e_s[1].item()),
                                  "h2d_bytes": int(h2d_first), "d2h_bytes": int(d2h_first),
                                  "phases_ms": {k: float(first_phases[-1][i] * 1e3) for i, k in enumerate(("create", "pattern", "assemble_to_host"))},
                                  "includes": "problem upload, 1-D tables, pattern build, assembly, outer/inner/values/rhs into pinned host memory"}}
        if e2e is None:
            e2e = e2e1
        else:
            e2e["first_assembly"] = e2e1["first_assembly"]
    res["e2e"] = e2e

    # ---------------- the reference's CPU assembly on this box's host cores (rank 0, N = 1)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import refutil as R
        mat = None
        if R.have_ref():     # in its own process: thread binding and thread count are the reference's, not this process's
            env = {k: v for k, v in os.environ.items() if k not in ("OMP_NUM_THREADS", "OMP_PROC_BIND")}
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--cpu-matrix", "--ref-nelem", str(args.ref_nelem)],
                                 capture_output=True, text=True, env=env, timeout=900)
            try:
                mat = json.loads(out.stdout.strip().splitlines()[-1])
            except Exception:
                mat = {"unavailable": (out.stderr or out.stdout)[-300:]}
        if mat and "gsPoissonAssembler_all" in mat:
            head = mat["gsPoissonAssembler_all"]
            cpu_baseline = {"value": head["value"], "unit": "DOFs/s", "cores": head["cores"], "kind": "reference",
                            "sample": "gsPoissonAssembler::assemble, " + head["sample"] + f", {head['seconds']:.2f} s, OMP_PROC_BIND=close",
                            "all_paths": mat}
        else:
            pbs = g.host.poisson_box_problem(3, 3, 12, g.expr_compile(F3))
            t0 = time.time(); R.oracle_assemble(pbs); dt = time.time() - t0
            cpu_baseline = {"value": pbs.nfree / dt, "unit": "DOFs/s", "cores": 1, "kind": "port",
                            "sample": f"oracle/gsb_oracle.c, 3D p=3, 12^3 elements, {pbs.nfree} DOFs, {dt:.2f} s", "reference": mat}
    res["cpu_baseline"] = cpu_baseline

    # ---------------- short runs of the multi-patch configs (the collective path), appended to the headline line
    if args.config == "2" and not args.no_extra:
        extra = {}
        for cfg in ("3", "4"):
            try:
                r, _ = run_config(cfg, 3, 3, full=False)
                extra[cfg] = {k: r[k] for k in ("value", "unit", "ms_per_step", "scaling", "nccl_bytes_per_step", "verification", "setup_ms", "gpu_launches")}
                extra[cfg]["workload"] = r["config"]["workload"]; extra[cfg]["parallelism"] = r["config"]["parallelism"]
            except Exception as e:  # noqa
                extra[cfg] = {"error": str(e)[:300]}
        res["configs"] = extra

    if world > 1:
        dist.destroy_process_group()
    sys.stdout.flush()
    os.dup2(real_stdout, 1)
    if rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
