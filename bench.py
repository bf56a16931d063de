#!/usr/bin/env python
"""Benchmark of the BIEM hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (config C3 of BASELINE.json / SURVEY 8d): 3-D ('ba') 4x4 grid of 16 unit spheres, n_end = 16 (H = 256,
N = 4096 unknowns per system), a 256-point wavenumber sweep k_i = 0.5 + 7.5 i / 255, plane wave along x0, eta = 1,
sound-soft; per wavenumber: right-hand side, assembly of the 4096 x 4096 complex128 system, blocked LU solve, and
u_scat at the origin + 64 probe points.  One STEP = the whole 256-system sweep; with N GPUs wavenumber i goes to rank
i mod N (no data-path collective; total work fixed => "scaling": "strong").

`value`  : systems/s with inputs resident in HBM (torch CUDA tensors in, CUDA tensors out).
`e2e`    : systems/s through the reference-shaped public API with HOST (NumPy) inputs and outputs: the H2D copies
           of the inputs and the D2H copies of density + u_scat are inside the timed region.  `e2e.c5` is the second
           half of the metric measured the same way (see below).
`roofline`: the LU trailing update (zgemm3m_tma_kernel, FP64 DMMA) -- flops of its launches / their summed CUDA-event
           durations inside one eager pass of a sweep group (the launches the sweep's graphs replay: `batch` systems per
           bhs_zgesv_batched call), against the FP64 tensor peak measured in the same run
           (MEASURED_PEAKS.json has no FP64 entry; the cuBLAS ZGEMM 8192^3 cross-check is printed beside it).
           `roofline.uscat`, `roofline.c5_lu` and `roofline.assembly` carry the other kernels of the metric.
`c5`     : config C5 end to end through the public API (default on): biem(keep_matrix=False) of the 64-sphere, n_end = 24
           system (N = 36 864 unknowns, 21.7 GB matrix) on rank 0 -> NCCL broadcast of the density -> every rank
           evaluates its row tile of the 2048^2 field grid with BIEMResultCalculator.uscat, pinned host grid in / pinned
           host field out.  `uscat` keeps the kernel-only points/s (real density, device-resident points).
`gpu_library_baseline`: the reference's own GPU path, torch.linalg.solve (cuSOLVER), on the same box and sizes.
`cpu_baseline` / `--impl reference`: the NumPy/SciPy oracle (reference algorithm, LAPACK zgesv) on the host cores, with
           its assembly / zgesv split and the BLAS thread count.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # one hardware queue per in-flight group of systems (see _biem._sweep_shape)

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "3D 16-sphere k-sweep systems/sec (assemble+solve)"  # first half of BASELINE.json's metric; `uscat` / `c5` carry the second
N_END = 16
HALF = 2
N_SYSTEMS = 256
N_PROBE = 64


# --------------------------------------------------------------------------------------------------------
def _args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--systems", type=int, default=N_SYSTEMS, help="wavenumbers in the sweep (default: the C3 256)")
    ap.add_argument("--no-uscat-grid", "--no-c5", dest="no_c5", action="store_true",
                    help="skip the C5 leg (N = 36 864 solve + 2048^2 field map)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-library-baseline", action="store_true")
    ap.add_argument("--cpu-systems", type=int, default=3, help="systems timed by the cpu_baseline leg")
    ap.add_argument("--grid", type=int, default=2048)
    ap.add_argument("--c5-n-end", type=int, default=24)
    ap.add_argument("--c5-half", type=int, default=4, help="the C5 geometry is a (2 half) x (2 half) grid of spheres")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                 str(self.index)], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        try:
            rows = [ln.strip().split(", ") for ln in open(self.path) if ln.strip()]
            os.unlink(self.path)
        except Exception:
            return out
        sm, mx, power = [], [], []
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for nm, v in zip(names, r[5:9]):
                if v.strip().lower().startswith("active"):
                    reasons.add(nm)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), power_w_max=float(max(power)),
                       reasons=sorted(reasons), samples=len(sm))
        return out


# --------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the NumPy/SciPy oracle on the host cores
# --------------------------------------------------------------------------------------------------------
def _blas_threads() -> int | None:
    try:
        from threadpoolctl import threadpool_info

        n = [p.get("num_threads") for p in threadpool_info() if p.get("user_api") == "blas"]
        return int(max(n)) if n else None
    except Exception:
        return None


def _oracle_systems(ks, with_probe=True):
    """Right-hand side + assembly + zgesv (+ probe u_scat) of the C3 system for each k with the oracle, the steps of
    oracle.biem spelled out so that they can be timed separately.  Returns (total seconds, last field, split)."""
    from oracle import biem_oracle as O

    from biem_helmholtz_sphere_b200.geometry import probe_ring

    cen = O.grid_centers(HALF, 3)
    B = cen.shape[0]
    rad = np.ones(B)
    al, be = np.ones(B, dtype=np.complex128), np.zeros(B, dtype=np.complex128)
    x = probe_ring(N_PROBE, 10.0, 3)
    split = {"rhs_s": 0.0, "assemble_s": 0.0, "zgesv_s": 0.0, "uscat_s": 0.0}
    t_all = time.perf_counter()
    last = None
    for k in ks:
        k = float(k)
        t0 = time.perf_counter()
        uin, _ = O.plane_wave(k=k, direction=np.array([1.0, 0.0, 0.0]))
        g = O.boundary_data("ba", cen, rad, N_END, al, be, uin, None)
        f_hat = O.expand("ba", g, N_END)
        t1 = time.perf_counter()
        A = O.assemble("ba", cen, rad, k, N_END, 1.0, al, be)
        t2 = time.perf_counter()
        H = f_hat.shape[1]
        dens = np.linalg.solve(A.reshape(B * H, B * H), f_hat.reshape(B * H)).reshape(B, H)
        t3 = time.perf_counter()
        if with_probe:
            res = O.OracleResult(c=O.OracleCoordinates("ba"), centers=cen.T.copy(), radii=rad, k=k, n_end=N_END, eta=1.0,
                                 kind="outer", density=dens, matrix=None)
            last = res.uscat(x)
        t4 = time.perf_counter()
        split["rhs_s"] += t1 - t0
        split["assemble_s"] += t2 - t1
        split["zgesv_s"] += t3 - t2
        split["uscat_s"] += t4 - t3
    return time.perf_counter() - t_all, last, split


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import biem_oracle as O

    from biem_helmholtz_sphere_b200.geometry import sweep_wavenumbers

    O.coupling_matrix("ba", N_END)  # k-independent table: built once, outside the timed region (as in our arm)
    ks = sweep_wavenumbers(args.systems)
    per_step = 1  # bounded sample: one system of the sweep per step (~3 s on 16 cores)
    pick = lambda i: ks[(i * 37) % len(ks)]  # noqa: E731  spread the samples over the sweep
    it = 0
    for _ in range(args.warmup):
        _oracle_systems([pick(it)])
        it += 1
    t = 0.0
    split = {}
    for _ in range(args.steps):
        dt, _, sp = _oracle_systems([pick(it + j) for j in range(per_step)])
        t += dt
        for kk, v in sp.items():
            split[kk] = split.get(kk, 0.0) + v
        it += per_step
    nsys = args.steps * per_step
    value = nsys / t
    cores = os.cpu_count()
    sample = (f"{per_step} system(s) of the {args.systems}-k sweep per step ({nsys} systems timed after {args.warmup} warm-up "
              f"systems), oracle (NumPy/SciPy restatement of the reference algorithm, LAPACK zgesv)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.systems, args.gpus),
        "cpu_baseline": {"value": value, "unit": "systems/s", "cores": cores, "kind": "port", "sample": sample,
                         "blas_threads": _blas_threads(),
                         "split_s_per_system": {kk: v / nsys for kk, v in split.items()},
                         "note": "one CPU process whatever --gpus says; split_s_per_system shows how a system's time divides between "
                                 "the assembly of the (S|R) blocks and LAPACK zgesv on this host"},
        "e2e": {"value": value, "unit": "systems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _config(nsys, gpus):
    return {
        "workload": f"C3: 3-D 'ba' 4x4 grid of 16 unit spheres (spacing 4), n_end=16 (N=4096 unknowns/system), "
                    f"{nsys}-point wavenumber sweep k=0.5..8, plane wave e0, eta=1, sound-soft; per k: rhs + assemble + "
                    f"LU solve + u_scat at origin and {N_PROBE} probe points",
        "systems_per_step": nsys, "n_unknowns": 4096, "sharding": f"k_i -> rank i mod {gpus}",
        "l2_policy": "inputs larger than L2 (each system streams its own 268 MB matrix; up to 64 in flight = 17 GB working set >> 126 MB L2)",
    }


def _load_profile(name: str) -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "profiles", name)))
    except Exception:
        return {}


def _measured_peaks() -> dict:
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


# --------------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    import biem_helmholtz_sphere_b200 as bhs
    from biem_helmholtz_sphere_b200 import _ops
    from biem_helmholtz_sphere_b200.geometry import grid_centers, probe_ring, sweep_wavenumbers

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F64, C128 = torch.float64, torch.complex128

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c = bhs.create_from_branching_types("ba")
    cen_np = grid_centers(HALF, 3)
    B = cen_np.shape[0]
    rad_np = np.ones(B)
    ks_all = sweep_wavenumbers(args.systems)
    ks_np = np.ascontiguousarray(ks_all[rank::world])
    K = len(ks_np)
    x_np = probe_ring(N_PROBE, 10.0, 3)
    dir_np = np.array([[1.0], [0.0], [0.0]])
    eta_np = np.ones(K)
    H = N_END * N_END
    N = B * H

    # ---- device-resident step --------------------------------------------------------------------------
    cen_d = torch.as_tensor(cen_np, device=dev)
    rad_d = torch.as_tensor(rad_np, device=dev)
    ks_d = torch.as_tensor(ks_np, device=dev)
    eta_d = torch.as_tensor(eta_np, device=dev)
    dir_d = torch.as_tensor(dir_np, device=dev)
    x_d = torch.as_tensor(x_np, device=dev)

    def step_resident():
        uin, _ = bhs.plane_wave(k=ks_d, direction=dir_d)
        res = bhs.biem(c, centers=cen_d[None], radii=rad_d[None], k=ks_d, n_end=N_END, eta=eta_d, uin=uin, keep_matrix=False)
        u = res.uscat(x_d)
        return res.density, u

    def step_host():
        uin, _ = bhs.plane_wave(k=ks_np, direction=dir_np)
        res = bhs.biem(c, centers=cen_np[None], radii=rad_np[None], k=ks_np, n_end=N_END, eta=eta_np, uin=uin, keep_matrix=False)
        u = res.uscat(x_np)
        return res.density, u

    def timed(fn, steps, sampler=None):
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        out = None
        for _ in range(steps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        clocks = sampler.stop() if sampler else None
        barrier()
        ms = max(e0.elapsed_time(e1), 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=F64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), out, clocks

    # warm-up (also builds the plan tables, captures the per-slot CUDA graphs)
    for _ in range(max(args.warmup, 3)):
        dens, u = step_resident()
    torch.cuda.synchronize()
    # launches of the pieces of a sweep, counted on eager calls, for gpu_launches: the right-hand sides of all K systems are one
    # launch group per step, assembly + LU one group per `batch` systems (bhs_zgesv_batched), the probe field one per system
    _ops.launch_count(reset=True)
    f1 = _ops.rhs_expand(3, N_END, centers=cen_d, radii=rad_d, k_in=ks_d[:1], direction=dir_d.reshape(3))
    torch.cuda.synchronize()
    launches_rhs = _ops.launch_count(reset=True)
    A2 = _ops.assemble(3, N_END, cen_d, rad_d, ks_d[:1].repeat(2), eta_d[:1].repeat(2))
    rhs1 = f1.reshape(N).clone()
    bufs2 = _ops.SolveBuffers(N, 1, 2)
    _ops.zgesv_batched_(A2, rhs1.repeat(2, 1), bufs2)
    torch.cuda.synchronize()
    launches_per_group = _ops.launch_count(reset=True)
    _ops.uscat(3, N_END, cen_d, rad_d, float(ks_np[0]), 1.0, rhs1.reshape(B, H), x_d)
    torch.cuda.synchronize()
    launches_uscat = _ops.launch_count(reset=True)
    del A2, bufs2

    from biem_helmholtz_sphere_b200._biem import _sweep_shape

    sweep_batch, sweep_groups = _sweep_shape(N, K)
    sampler = ClockSampler(local) if rank == 0 else None
    ms, wall_ms, (dens, u), clocks = timed(step_resident, args.steps, sampler)
    value = args.systems * args.steps / (ms * 1e-3)

    # ---- e2e: host arrays in, host arrays out ----------------------------------------------------------
    step_host()
    ms_h, wall_h, (dens_h, u_h), _ = timed(step_host, args.steps)
    # host-side wall clock is the honest figure here (the D2H copies synchronise)
    e2e_ms = max(ms_h, wall_h)
    e2e_value = args.systems * args.steps / (e2e_ms * 1e-3)
    h2d = int(ks_np.nbytes + eta_np.nbytes + cen_np.nbytes + rad_np.nbytes + dir_np.nbytes + x_np.nbytes + ks_np.nbytes)
    d2h = int(np.asarray(dens_h).nbytes + np.asarray(u_h).nbytes)

    # sanity: resident and host paths agree, nothing is NaN
    dd = dens.detach().cpu().numpy()
    if not (os.environ.get("BHS_LU_SKIP") or os.environ.get("BHS_LU_GEMM_ONLY")):  # (measurement aids: wrong results by design)
        assert np.all(np.isfinite(dd)) and np.all(np.isfinite(np.asarray(u_h)))
        assert np.allclose(dd, np.asarray(dens_h), rtol=1e-12, atol=1e-14)

    out = {
        "metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(args.systems, world),
        "e2e": {"value": e2e_value, "unit": "systems/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int((launches_rhs + launches_per_group * (-(-K // sweep_batch)) + launches_uscat * K) * args.steps * world),
        "launches_per_group": int(launches_per_group), "systems_per_launch": int(sweep_batch),
        "clocks": clocks,
    }

    # ---- per-kernel split + roofline of the LU trailing update (rank 0, eager profiled pass) -------------
    if rank == 0:
        traffic = _load_profile("r02c_zgemm_traffic.json") or _load_profile("r02_zgemm_traffic.json")
        peaks = {"dfma": _ops.fp64_peak(0, 4096), "dmma884": _ops.fp64_peak(1, 4096), "dmma1684": _ops.fp64_peak(2, 2048),
                 "dmma1688": _ops.fp64_peak(3, 1024), "dmma16816": _ops.fp64_peak(4, 512)}
        # cross-check of the tensor peak with the vendor library: cuBLAS ZGEMM 8192^3 (8 real flops per complex FMA)
        ga = torch.randn(8192, 8192, dtype=C128, device=dev)
        gb = torch.randn(8192, 8192, dtype=C128, device=dev)
        torch.matmul(ga, gb)
        z0, z1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        z0.record()
        torch.matmul(ga, gb)
        z1.record()
        torch.cuda.synchronize()
        peaks["cublas_zgemm_8192"] = 8.0 * 8192.0 ** 3 / (z0.elapsed_time(z1) * 1e-3) * 1e-12
        del ga, gb
        peaks["note"] = ("TFLOP/s; dfma / dmma*: register-resident loops of bhs_fp64_peak (every mma.sync f64 shape lowers to "
                         "DMMA.8x8x4 on sm_100a, profiles/r02_fp64_mma_shapes.txt); cublas_zgemm_8192: torch.matmul complex128")
        peak_dmma = max(peaks["dmma884"], 1e-9)
        peak_dfma = max(peaks["dfma"], 1e-9)
        # Profiled pass = what the sweep's graphs replay: ONE group of `sweep_batch` systems through bhs_assemble (nsys = batch)
        # and bhs_zgesv_batched, launched eagerly with an event pair around every launch (bhs_profile).  A second pass times a
        # lone system through bhs_zgesv (cluster panels + look-ahead), for kernel_split_lone_ms.
        nprof = max(1, min(sweep_batch, K))
        A = torch.empty((nprof, N, N), dtype=C128, device=dev)
        work = _ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3, N_END).handle, B, nprof))
        bufs_b = _ops.SolveBuffers(N, 1, nprof)
        bufs = _ops.SolveBuffers(N, 1)

        def group_pass():
            f = _ops.rhs_expand(3, N_END, centers=cen_d, radii=rad_d, k_in=ks_d[:nprof], direction=dir_d.reshape(3))
            _ops.assemble(3, N_END, cen_d, rad_d, ks_d[:nprof], eta_d[:nprof], out=A, work=work)
            r = f.reshape(nprof, N).clone()
            _ops.zgesv_batched_(A, r, bufs_b)
            for i in range(nprof):
                _ops.uscat(3, N_END, cen_d, rad_d, float(ks_np[i]), 1.0, r[i].reshape(B, H), x_d)
            return f

        def lone_pass():
            f = _ops.rhs_expand(3, N_END, centers=cen_d, radii=rad_d, k_in=ks_d[:1], direction=dir_d.reshape(3))
            _ops.assemble(3, N_END, cen_d, rad_d, ks_d[:1], eta_d[:1], out=A[:1], work=work)
            r = f.reshape(N).clone()
            _ops.zgesv_(A[0], r, bufs)
            _ops.uscat(3, N_END, cen_d, rad_d, float(ks_np[0]), 1.0, r.reshape(B, H), x_d)
            return f

        group_pass()  # untimed: loads every kernel of the batched path
        torch.cuda.synchronize()
        _ops.profile(True)
        lone_pass()
        torch.cuda.synchronize()
        prof_lone = _ops.profile_read()
        _ops.profile(False)
        _ops.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f = group_pass()
        e1.record()
        torch.cuda.synchronize()
        prof = _ops.profile_read()
        _ops.profile(False)
        f = f[:1]
        tot_ms = e0.elapsed_time(e1)
        g = prof["lu_gemm"]
        gemm_tf = g["work"] / (g["ms"] * 1e-3) * 1e-12 if g["ms"] > 0 else 0.0
        lu_ms = sum(prof[n]["ms"] for n in ("lu_gemm", "lu_gemm_inner", "lu_panel", "lu_trsm", "lu_pack", "lu_rhs"))
        gi = prof["lu_gemm_inner"]
        am = prof["asm_main"]
        out["roofline"] = {
            "bound": "tensor", "kernel": "zgemm3m_tma_kernel (LU trailing update, FP64 DMMA, operands by tensor-map TMA, three real "
                                         "products per complex one)",
            "achieved": gemm_tf, "peak": peak_dmma, "unit": "TFLOP/s", "frac": gemm_tf / peak_dmma,
            "executed_tflops": 0.75 * gemm_tf, "executed_frac": 0.75 * gemm_tf / peak_dmma,
            "flop_convention": "achieved / frac count the ALGORITHMIC flops of SURVEY 8d (8 real flops per complex multiply-add, "
                               "(8/3) N^3 per factorisation); the kernel uses the 3M scheme (T1 = Ar Br, T2 = Ai Bi, T3 = (Ar+Ai)(Br+Bi)) "
                               "and EXECUTES 6, so the tensor pipe itself is busy at executed_frac",
            "peak_source": "FP64 mma.sync m8n8k4 register-resident loop measured in this run (bhs_fp64_peak); "
                           "MEASURED_PEAKS.json carries no FP64 figure; cuBLAS ZGEMM 8192^3 in this run: "
                           f"{peaks['cublas_zgemm_8192']:.2f} TFLOP/s",
            "traffic": traffic.get("traffic_bytes"), "traffic_note": traffic.get("launch", "no ncu capture found")
            + " -- dram__bytes_read.sum + dram__bytes_write.sum of that one launch (algorithmic bytes of the same launch: "
            + str(traffic.get("algorithmic_bytes")) + ")",
            "launches": int(g["count"]), "avg_launch_ms": g["ms"] / max(g["count"], 1),
            "flops_per_launch_avg": g["work"] / max(g["count"], 1),
            "scope": f"every update with K >= 128 of one group of {nprof} systems factorised in lock step by bhs_zgesv_batched -- the "
                     "launches the sweep's CUDA graphs replay (256-wide outer blocks: K = 256 trailing updates, K = 128 updates "
                     "inside a block); the K = 32/64 updates inside the panel recursion are reported under inner_updates",
            "inner_updates": {"tflops": gi["work"] / (gi["ms"] * 1e-3) * 1e-12 if gi["ms"] > 0 else None,
                              "launches": int(gi["count"]), "ms_per_system": gi["ms"] / nprof},
            "fp64_peaks": peaks,
        }
        # the same kernel on the largest trailing-update shape of a C3 factorisation, launched back to back (steady clocks)
        Mt = N - 128
        La = torch.randn(Mt, 128, dtype=C128, device=dev)
        Ua = torch.randn(128, Mt, dtype=C128, device=dev)
        Ca = torch.randn(Mt, Mt, dtype=C128, device=dev)
        wk = _ops._work(_ops.load().bhs_zgemm_workspace(Mt, Mt, 128))
        for _ in range(3):
            _ops.zgemm_sub_(Ca, La, Ua, work=wk)
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(20):
            _ops.zgemm_sub_(Ca, La, Ua, work=wk)
        g1.record()
        torch.cuda.synchronize()
        iso_tf = 20 * 8.0 * Mt * Mt * 128 / (g0.elapsed_time(g1) * 1e-3) * 1e-12
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.addmm(Ca, La, Ua, alpha=-1.0, out=Ca)
        c0.record()
        for _ in range(20):
            torch.addmm(Ca, La, Ua, alpha=-1.0, out=Ca)
        c1.record()
        torch.cuda.synchronize()
        cub_tf = 20 * 8.0 * Mt * Mt * 128 / (c0.elapsed_time(c1) * 1e-3) * 1e-12
        out["roofline"]["isolated_largest_launch"] = {"shape": f"{Mt}x{Mt}x128", "tflops": iso_tf, "frac": iso_tf / peak_dmma,
                                                      "cublas_zgemm_same_shape_tflops": cub_tf}
        del La, Ua, Ca, wk
        out["kernel_split_ms_per_system"] = {n: v["ms"] / nprof for n, v in prof.items()}
        out["kernel_split_ms_per_system"]["eager_total"] = tot_ms / nprof
        out["kernel_split_note"] = (f"one group of {nprof} systems in lock step (bhs_assemble nsys = {nprof}, bhs_zgesv_batched), eager "
                                    "launches, per system; kernel_split_lone_ms: ONE lone system through bhs_zgesv (cluster panels, look-ahead)")
        out["kernel_split_lone_ms"] = {n: v["ms"] for n, v in prof_lone.items()}
        lu_ms = sum(prof_lone[n]["ms"] for n in ("lu_gemm", "lu_gemm_inner", "lu_panel", "lu_trsm", "lu_pack", "lu_rhs"))
        # a lone system as a CUDA graph (what the sweep engine replays), timed alone
        Ag = torch.empty((N, N), dtype=C128, device=dev)
        _ops.assemble(3, N_END, cen_d, rad_d, ks_d[:1], eta_d[:1], out=A[:1], work=work)
        rg = f.reshape(N).clone()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            Ag.copy_(A[0])
            _ops.zgesv_(Ag, rg, bufs)
            side.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=side):
                _ops.zgesv_(Ag, rg, bufs)
            tg = 0.0
            for _ in range(5):
                Ag.copy_(A[0])
                q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                q0.record(side)
                gr.replay()
                q1.record(side)
                side.synchronize()
                tg += q0.elapsed_time(q1)
        lu_graph_ms = tg / 5
        flops_lu = (8.0 / 3.0) * N ** 3
        out["lu"] = {"tflops": flops_lu / (lu_graph_ms * 1e-3) * 1e-12, "ms": lu_graph_ms,
                     "eager_ms": lu_ms, "eager_tflops": flops_lu / (lu_ms * 1e-3) * 1e-12 if lu_ms > 0 else None,
                     "frac_of_dmma_peak": flops_lu / (lu_graph_ms * 1e-3) * 1e-12 / peak_dmma,
                     "what": "ONE N = 4096 system factorised and solved alone (bhs_zgesv as a CUDA graph; eager_*: plain launches)",
                     "fp64_dmma_peak_tflops": peak_dmma, "fp64_dfma_peak_tflops": peak_dfma}
        del Ag, gr
        # ---- the reference's own GPU path on this box: torch.linalg.solve -> cuSOLVER (_biem.py:797 with torch CUDA tensors) ----
        if not args.no_library_baseline:
            out["gpu_library_baseline"] = _library_baseline(torch, dev, N, A[0], f.reshape(N), flops_lu)
        hbm = _measured_peaks().get("hbm_gbs", 6650.0)
        # the assembly kernel only WRITES (16 N^2 bytes): also measure the pure write stream of this device (fill of 4 GiB)
        wbuf = torch.empty(1 << 30, dtype=torch.int32, device=dev)
        wbest = 0.0
        for _ in range(6):
            w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            w0.record()
            wbuf.fill_(1)
            w1.record()
            torch.cuda.synchronize()
            wbest = max(wbest, wbuf.numel() * 4 / (w0.elapsed_time(w1) * 1e-3) * 1e-9)
        del wbuf
        if am["ms"] > 0:
            gbs = am["work"] / (am["ms"] * 1e-3) * 1e-9
            out["assembly"] = {"gbs": gbs, "ms": am["ms"] / nprof, "hbm_peak_gbs": hbm, "frac_hbm": gbs / hbm,
                               "write_stream_gbs": wbest, "frac_write_stream": gbs / wbest,
                               "note": "algorithmic bytes = 16 N^2 per system (SURVEY 8d); hbm_peak_gbs is MEASURED_PEAKS.json's copy "
                                       "figure (read + write bytes); write_stream_gbs is a 4 GiB fill measured in this run"}
            out["roofline"]["assembly"] = {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm}
        out["special_functions"] = _bench_special(torch, _ops, dev, hbm)
        del A, work, bufs, bufs_b

    # ---- C5: the 36 864-unknown system + the 2048^2 field map through the public API, field rows split over the ranks ----
    if not args.no_c5:
        from biem_helmholtz_sphere_b200._biem import clear_engines

        clear_engines()  # the C3 slot buffers (17 GB) are not needed any more
        torch.cuda.empty_cache()
        c5 = _bench_c5(args, bhs, _ops, torch, dev, rank, world, barrier, dist if world > 1 else None)
        if rank == 0:
            out["c5"] = c5["c5"]
            out["uscat"] = c5["uscat"]
            out["e2e"]["c5"] = c5["e2e"]
            if "roofline" in out:
                out["roofline"]["uscat"] = c5["uscat"]["roofline"]
                out["roofline"]["c5_lu"] = c5["c5"].get("lu_roofline")

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import biem_oracle as O

        O.coupling_matrix("ba", N_END)
        ncpu = max(1, args.cpu_systems)
        sel = [ks_all[(i * 37) % len(ks_all)] for i in range(ncpu)]
        dt, _, sp = _oracle_systems(sel)
        out["cpu_baseline"] = {
            "value": ncpu / dt, "unit": "systems/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{ncpu} systems of the sweep (k = {', '.join(f'{v:.3f}' for v in sel)}), NumPy/SciPy oracle, "
                      f"LAPACK zgesv with all host cores",
            "blas_threads": _blas_threads(), "split_s_per_system": {kk: v / ncpu for kk, v in sp.items()},
        }
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _library_baseline(torch, dev, N, A, f, flops_lu) -> dict:
    """torch.linalg.solve (cuSOLVER) on the C3 system: one system, a batch of 8, and 8 streams of single solves."""
    def ev(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # MAGMA prints a banner to the C-level stdout for large batched solves: keep it out of this program's single JSON line
    sys.stdout.flush()
    saved_fd = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 1)
    try:
        return _library_baseline_inner(torch, ev, N, A, f, flops_lu)
    finally:
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
        os.close(devnull)


def _library_baseline_inner(torch, ev, N, A, f, flops_lu) -> dict:
    b = f.reshape(N, 1).clone()
    ms1 = ev(lambda: torch.linalg.solve(A, b), 5)
    A8 = A.expand(8, N, N).contiguous()
    b8 = b.expand(8, N, 1).contiguous()
    ms8 = ev(lambda: torch.linalg.solve(A8, b8), 2)
    streams = [torch.cuda.Stream() for _ in range(8)]

    def multi():
        for i, s in enumerate(streams):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                torch.linalg.solve(A8[i], b8[i])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)

    msm = ev(multi, 2)
    return {
        "what": "the reference's own GPU path: torch.linalg.solve complex128 -> cuSOLVER (_biem.py:797 with torch CUDA tensors), "
                "same box, same N = 4096 system",
        "single_ms": ms1, "single_tflops": flops_lu / (ms1 * 1e-3) * 1e-12,
        "batched8_systems_per_s": 8.0 / (ms8 * 1e-3), "streams8_systems_per_s": 8.0 / (msm * 1e-3),
    }


def _bench_special(torch, _ops, dev, hbm) -> dict:
    """K1 / K2 standalone (HBM-store bound, 16 B per (argument, order) / (direction, harmonic))."""
    def ev(fn, reps=5):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    g = torch.Generator(device=dev).manual_seed(1)
    nx, n_max = 1 << 21, 31
    x = torch.rand(nx, dtype=torch.float64, device=dev, generator=g) * 40.0 + 0.5
    ms_b = ev(lambda: _ops.bessel(3, 2, n_max, x))
    gb_b = nx * (n_max + 1) * 16 / (ms_b * 1e-3) * 1e-9
    npts, n_end = 1 << 18, 16
    xyz = torch.randn(3, npts, dtype=torch.float64, device=dev, generator=g)
    ms_h = ev(lambda: _ops.harmonics(3, n_end, xyz))
    gb_h = npts * n_end * n_end * 16 / (ms_h * 1e-3) * 1e-9
    return {
        "k1_bessel": {"what": f"h_n^(1), n = 0..{n_max}, {nx} real arguments (3-D)", "ms": ms_b, "gbs": gb_b, "frac_hbm": gb_b / hbm},
        "k2_harmonics": {"what": f"Y_h, n_end = {n_end} (H = {n_end * n_end}), {npts} directions (3-D)", "ms": ms_h, "gbs": gb_h,
                         "frac_hbm": gb_h / hbm},
        "note": "algorithmic bytes = 16 B per stored value (SURVEY 8d), against MEASURED_PEAKS.json's HBM figure",
    }


def _bench_c5(args, bhs, _ops, torch, dev, rank, world, barrier, dist):
    """Config C5 through the public API: rank 0 assembles and solves the (2 half)^2-sphere system (default 64 spheres,
    n_end = 24: N = 36 864 unknowns, 21.7 GB matrix) with biem(keep_matrix=False); the density is broadcast over NCCL (the one
    collective of the path); every rank evaluates its contiguous row tile of the G x G field grid on x2 = 0 through
    BIEMResultCalculator.uscat with pinned host points in and a pinned host field out."""
    from biem_helmholtz_sphere_b200 import parallel
    from biem_helmholtz_sphere_b200._biem import BIEMResultCalculator, clear_engines
    from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers

    n_end, half, k = args.c5_n_end, args.c5_half, 1.0
    c = bhs.create_from_branching_types("ba")
    cen_np = grid_centers(half, 3)
    B = cen_np.shape[0]
    rad_np = np.ones(B)
    H = n_end * n_end
    N = B * H
    G = args.grid
    F64, C128 = torch.float64, torch.complex128
    info: dict = {}
    dens_np = None
    t_pipe0 = time.perf_counter()
    if rank == 0:
        _ops.get_plan(3, n_end)  # k-independent tables (as for C3: outside the timed solve)
        torch.cuda.synchronize()
        uin, _ = bhs.plane_wave(k=np.asarray(k), direction=np.array([1.0, 0.0, 0.0]))
        _ops.profile(True)
        t0 = time.perf_counter()
        res0 = bhs.biem(c, centers=cen_np, radii=rad_np, k=np.asarray(k), n_end=n_end, eta=np.asarray(1.0), uin=uin,
                        keep_matrix=False)
        torch.cuda.synchronize()
        t_biem = time.perf_counter() - t0
        prof = _ops.profile_read()
        _ops.profile(False)
        dens_np = np.asarray(res0.density)
        lu_ms = sum(prof[n]["ms"] for n in ("lu_gemm", "lu_gemm_inner", "lu_panel", "lu_trsm", "lu_pack", "lu_rhs"))
        flops_lu = (8.0 / 3.0) * float(N) ** 3
        peak_dmma = max(_ops.fp64_peak(1, 4096), 1e-9)
        hbm = _measured_peaks().get("hbm_gbs", 6650.0)
        asm_ms = prof["asm_main"]["ms"]
        info = {
            "workload": f"C5: {B} unit spheres ({2 * half}x{2 * half} grid, spacing 4), n_end={n_end} (H={H}, N={N} unknowns, "
                        f"{16.0 * N * N / 1e9:.2f} GB matrix), k=1, plane wave e0, sound-soft; biem(keep_matrix=False) on rank 0",
            "biem_s": t_biem, "assemble_ms": asm_ms + prof["asm_pre"]["ms"], "assemble_main_ms": asm_ms,
            "assemble_gbs": 16.0 * N * N / (asm_ms * 1e-3) * 1e-9 if asm_ms > 0 else None,
            "assemble_frac_hbm": 16.0 * N * N / (asm_ms * 1e-3) * 1e-9 / hbm if asm_ms > 0 else None,
            "lu_ms": lu_ms, "lu_tflops": flops_lu / (lu_ms * 1e-3) * 1e-12 if lu_ms > 0 else None,
            "lu_frac": flops_lu / (lu_ms * 1e-3) * 1e-12 / peak_dmma if lu_ms > 0 else None,
            "lu_gemm_ms": prof["lu_gemm"]["ms"],
            "lu_gemm_tflops": prof["lu_gemm"]["work"] / (prof["lu_gemm"]["ms"] * 1e-3) * 1e-12 if prof["lu_gemm"]["ms"] > 0 else None,
            "lu_split_ms": {n: prof[n]["ms"] for n in ("lu_gemm", "lu_gemm_inner", "lu_panel", "lu_trsm", "lu_rhs")},
            "lu_roofline": {"bound": "tensor", "achieved": flops_lu / (lu_ms * 1e-3) * 1e-12 if lu_ms > 0 else None, "peak": peak_dmma,
                            "unit": "TFLOP/s", "frac": flops_lu / (lu_ms * 1e-3) * 1e-12 / peak_dmma if lu_ms > 0 else None,
                            "what": f"whole factorisation + solve of the N = {N} system ((8/3) N^3 flops), one GPU"},
        }
        clear_engines()  # frees the 21.7 GB slot
        torch.cuda.empty_cache()
        # size-independent check at full size: residual of the solved system against a fresh assembly
        try:
            cen_t = torch.as_tensor(cen_np, device=dev)
            rad_t = torch.as_tensor(rad_np, device=dev)
            kt = torch.tensor([k], dtype=F64, device=dev)
            A = _ops.assemble(3, n_end, cen_t, rad_t, kt, torch.ones(1, dtype=F64, device=dev))[0]
            f = _ops.rhs_expand(3, n_end, centers=cen_t, radii=rad_t, k_in=kt,
                                direction=torch.tensor([1.0, 0.0, 0.0], dtype=F64, device=dev)).reshape(N)
            x = torch.as_tensor(dens_np, device=dev).reshape(N)
            info["solve_rel_residual"] = float(torch.linalg.vector_norm(A @ x - f) / torch.linalg.vector_norm(f))
            del A, f, x
            torch.cuda.empty_cache()
        except Exception as e:  # pragma: no cover  (out of memory on a small device)
            info["solve_rel_residual"] = None
            info["residual_error"] = str(e)[:200]
    # ---- the one collective of the path: solved coefficients -> every GPU -------------------------------------------
    dens_t = parallel.broadcast_density(None if dens_np is None else torch.as_tensor(dens_np, device=dev), (B, H), src=0)
    if rank == 0:
        res = res0
    else:
        res = BIEMResultCalculator(c=c, centers=np.ascontiguousarray(cen_np.T), radii=rad_np, k=np.asarray(k), n_end=n_end,
                                   eta=np.asarray(1.0), kind="outer", density=dens_t.cpu().numpy())
    rows = parallel.field_rows(G, rank, world)
    x_np = np.ascontiguousarray(field_grid(G, 20.0, 3)[:, rows, :])  # [3, rows, G]
    P = x_np.shape[1] * x_np.shape[2]
    x_pin = torch.as_tensor(x_np).pin_memory()
    # ---- e2e: pinned host grid in -> BIEMResultCalculator.uscat -> pinned host field out ----------------------------
    for _ in range(3):  # warm-up (also lets torch's pinned-host allocator cache the two result buffers that alternate below)
        u = res.uscat(x_pin)
    reps = 5
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        u = res.uscat(x_pin)
    torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], dtype=F64, device=dev)
    barrier()
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    t_pipe = torch.tensor([time.perf_counter() - t_pipe0], dtype=F64, device=dev)
    if dist is not None:
        dist.all_reduce(t_pipe, op=dist.ReduceOp.MAX)
    assert isinstance(u, torch.Tensor) and u.is_pinned() and tuple(u.shape) == tuple(x_np.shape[1:])
    nan_frac = float(torch.isnan(u.real).double().mean())
    # ---- kernel only: device-resident points, real density ------------------------------------------------------------
    cen_t = torch.as_tensor(cen_np, device=dev)
    rad_t = torch.as_tensor(rad_np, device=dev)
    xd = x_pin.to(dev).reshape(3, -1).contiguous()
    work = _ops._work(_ops.load().bhs_uscat_workspace(_ops.get_plan(3, n_end).handle, B))
    for _ in range(2):
        o = _ops.uscat(3, n_end, cen_t, rad_t, k, 1.0, dens_t, xd, work=work)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        o = _ops.uscat(3, n_end, cen_t, rad_t, k, 1.0, dens_t, xd, work=work)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=F64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    # the general (non-coplanar) variant of the kernel beside it: the same tile lifted off the plane of the centres
    xg = xd.clone()
    xg[2] += 0.37
    for _ in range(2):
        _ops.uscat(3, n_end, cen_t, rad_t, k, 1.0, dens_t, xg, work=work)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(3):
        _ops.uscat(3, n_end, cen_t, rad_t, k, 1.0, dens_t, xg, work=work)
    g1.record()
    torch.cuda.synchronize()
    tg = torch.tensor([g0.elapsed_time(g1) / 3], dtype=F64, device=dev)
    if dist is not None:
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
    ms_gen = float(tg[0])
    peak_dfma = max(_ops.fp64_peak(0, 4096), 1e-9) if rank == 0 else 1.0
    Ptot = G * G
    flops = 8.0 * Ptot * B * H
    tf = flops / (ms * 1e-3) * 1e-12 / world
    tf_gen = flops / (ms_gen * 1e-3) * 1e-12 / world
    uscat = {
        "metric": "uscat points/sec (second half of BASELINE.json's metric)", "value": Ptot / (ms * 1e-3), "unit": "points/s", "ms": ms,
        "workload": f"C5: {B} unit spheres, n_end={n_end} (H={H}), k=1, {G}x{G} field grid on x2=0 over [-20,20]^2, "
                    f"rows split over {world} rank(s); density of the solved system, broadcast from rank 0",
        "roofline": {"bound": "fp64", "achieved": tf, "peak": peak_dfma, "unit": "TFLOP/s", "frac": tf / peak_dfma,
                     "points_per_s": Ptot / (ms * 1e-3), "n_gpus": world,
                     "general_variant": {"achieved": tf_gen, "frac": tf_gen / peak_dfma, "ms": ms_gen,
                                         "what": "same tile lifted 0.37 off the plane of the centres: the general kernel (Legendre "
                                                 "recurrence per (n, m), 8 FP64 instructions per step, all n (n + 1) / 2 pairs)"},
                     "note": "counted flops = 8 P B H (one complex FMA per point x ball x harmonic; special-function "
                             "generation not counted); per-GPU figure against the DFMA peak measured in this run. The grid and "
                             "the sphere centres are coplanar (x2 = 0), so the device-selected planar kernel runs: coefficients rotated "
                             "into the frame whose polar axis is the plane's normal, no Legendre recurrence, half of the (n, m) "
                             "pairs vanish -- it executes ~2.8x fewer FP64 instructions than the count above assumes, hence frac > 1; "
                             "general_variant is the kernel for arbitrary points"},
        "nan_fraction": nan_frac,
        "ncu": _load_profile("r02_uscat_ncu.json") or _load_profile("r01_uscat_ncu.json"),
    }
    e2e = {"points_per_s": Ptot / (float(te[0]) * 1e-3), "unit": "points/s", "ms": float(te[0]),
           "h2d_bytes": int(24 * Ptot), "d2h_bytes": int(16 * Ptot), "n_gpus": world,
           "pipeline_s": float(t_pipe[0]),
           "what": "BIEMResultCalculator.uscat per rank on its row tile, pinned host points in / pinned host field out, max over "
                   "ranks; pipeline_s = solve on rank 0 + density broadcast + warm-up and 5 timed maps"}
    info["uscat_points_s"] = uscat["value"]
    info["e2e_points_s"] = e2e["points_per_s"]
    return {"c5": info, "uscat": uscat, "e2e": e2e}


if __name__ == "__main__":
    a = _args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
