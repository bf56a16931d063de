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
           of the inputs and the D2H copies of density + u_scat are inside the timed region.
`roofline`: the LU trailing update (zgemm_sub_kernel, FP64 DMMA) -- flops of its launches / their summed CUDA-event
           durations inside one profiled sweep pass, against the FP64 tensor peak measured in the same run
           (MEASURED_PEAKS.json has no FP64 entry).
`uscat`  : the second half of the metric -- u_scat points/s on the 2048^2 grid of config C5 (64 spheres, n_end = 24).
`cpu_baseline` / `--impl reference`: the NumPy/SciPy oracle (reference algorithm, LAPACK zgesv) on the host cores.
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

METRIC = "3D 16-sphere k-sweep systems/sec (assemble+solve)"  # first half of BASELINE.json's metric; `uscat` carries the second
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
    ap.add_argument("--no-uscat-grid", action="store_true", help="skip the C5 2048^2 field-evaluation measurement")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-systems", type=int, default=3, help="systems timed by the cpu_baseline leg")
    ap.add_argument("--grid", type=int, default=2048)
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
def _oracle_systems(ks, with_probe=True):
    """Assemble + solve (+ probe u_scat) the C3 system for each k with the oracle; returns seconds."""
    from oracle import biem_oracle as O

    cen = O.grid_centers(HALF, 3)
    rad = np.ones(cen.shape[0])
    from biem_helmholtz_sphere_b200.geometry import probe_ring

    x = probe_ring(N_PROBE, 10.0, 3)
    t0 = time.perf_counter()
    last = None
    for k in ks:
        uin, _ = O.plane_wave(k=float(k), direction=np.array([1.0, 0.0, 0.0]))
        r = O.biem("ba", centers=cen, radii=rad, k=float(k), n_end=N_END, uin=uin, eta=1.0)
        if with_probe:
            last = r.uscat(x)
    return time.perf_counter() - t0, last


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import biem_oracle as O
    from biem_helmholtz_sphere_b200.geometry import sweep_wavenumbers

    O.coupling_matrix("ba", N_END)  # k-independent table: built once, outside the timed region (as in our arm)
    ks = sweep_wavenumbers(args.systems)
    per_step = 1  # bounded sample: one system of the sweep per step (~5 s on 8 cores)
    pick = lambda i: ks[(i * 37) % len(ks)]  # noqa: E731  spread the samples over the sweep
    it = 0
    warm_done = min(args.warmup, 1)  # one warm-up system is enough for a CPU loop; keeps the run bounded
    for _ in range(warm_done):
        _oracle_systems([pick(it)])
        it += 1
    t = 0.0
    for _ in range(args.steps):
        dt, _ = _oracle_systems([pick(it + j) for j in range(per_step)])
        t += dt
        it += per_step
    value = args.steps * per_step / t
    cores = os.cpu_count()
    sample = f"{per_step} system(s) of the {args.systems}-k sweep per step, oracle (NumPy/SciPy, LAPACK zgesv, BLAS threads = all cores)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm_done, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": _config(args.systems, args.gpus),
        "cpu_baseline": {"value": value, "unit": "systems/s", "cores": cores, "kind": "port", "sample": sample},
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


# --------------------------------------------------------------------------------------------------------
def run_b200(args) -> None:
    import torch
    import torch.distributed as dist

    import biem_helmholtz_sphere_b200 as bhs
    from biem_helmholtz_sphere_b200 import _ops
    from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers, probe_ring, sweep_wavenumbers

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
    # launches of one eager (assemble + solve + uscat) system, for gpu_launches
    _ops.launch_count(reset=True)
    A1 = _ops.assemble(3, N_END, cen_d, rad_d, ks_d[:1], eta_d[:1])
    f1 = _ops.rhs_expand(3, N_END, centers=cen_d, radii=rad_d, k_in=ks_d[:1], direction=dir_d.reshape(3))
    rhs1 = f1.reshape(N).clone()
    _ops.zgesv_(A1[0], rhs1)
    _ops.uscat(3, N_END, cen_d, rad_d, float(ks_np[0]), 1.0, rhs1.reshape(B, H), x_d)
    torch.cuda.synchronize()
    launches_per_system = _ops.launch_count(reset=True)
    del A1

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
    assert np.all(np.isfinite(dd)) and np.all(np.isfinite(np.asarray(u_h)))
    assert np.allclose(dd, np.asarray(dens_h), rtol=1e-12, atol=1e-14)

    out = {
        "metric": METRIC, "value": value, "unit": "systems/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": _config(args.systems, world),
        "e2e": {"value": e2e_value, "unit": "systems/s", "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
                "ms_per_step": e2e_ms / args.steps},
        # the sweep engine issues one launch of every kernel per GROUP of `batch` systems
        "gpu_launches": int(launches_per_system * (-(-K // sweep_batch)) * args.steps * world),
        "launches_per_system": int(launches_per_system), "systems_per_launch": int(sweep_batch),
        "clocks": clocks,
    }

    # ---- per-kernel split + roofline of the LU trailing update (rank 0, eager profiled pass) -------------
    if rank == 0:
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "r01_zgemm_traffic.json")))
        except Exception:
            traffic = {}
        peak_dmma = max(_ops.fp64_peak(1, 4096), 1e-9)
        peak_dfma = max(_ops.fp64_peak(0, 4096), 1e-9)
        nprof = min(4, K)
        A = torch.empty((1, N, N), dtype=C128, device=dev)
        work = _ops._work(_ops.load().bhs_assemble_workspace(_ops.get_plan(3, N_END).handle, B, 1))
        bufs = _ops.SolveBuffers(N, 1)
        torch.cuda.synchronize()
        _ops.profile(True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(nprof):
            f = _ops.rhs_expand(3, N_END, centers=cen_d, radii=rad_d, k_in=ks_d[i:i + 1], direction=dir_d.reshape(3))
            _ops.assemble(3, N_END, cen_d, rad_d, ks_d[i:i + 1], eta_d[i:i + 1], out=A, work=work)
            r = f.reshape(N).clone()
            _ops.zgesv_(A[0], r, bufs)
            _ops.uscat(3, N_END, cen_d, rad_d, float(ks_np[i]), 1.0, r.reshape(B, H), x_d)
        e1.record()
        torch.cuda.synchronize()
        prof = _ops.profile_read()
        _ops.profile(False)
        tot_ms = e0.elapsed_time(e1)
        g = prof["lu_gemm"]
        gemm_tf = g["work"] / (g["ms"] * 1e-3) * 1e-12 if g["ms"] > 0 else 0.0
        lu_ms = sum(prof[n]["ms"] for n in ("lu_gemm", "lu_gemm_inner", "lu_panel", "lu_trsm", "lu_pack", "lu_rhs"))
        gi = prof["lu_gemm_inner"]
        am = prof["asm_main"]
        out["roofline"] = {
            "bound": "tensor", "kernel": "zgemm_sub_kernel (LU trailing update, FP64 DMMA)",
            "achieved": gemm_tf, "peak": peak_dmma, "unit": "TFLOP/s", "frac": gemm_tf / peak_dmma,
            "peak_source": "FP64 mma.sync m8n8k4 register-resident loop measured in this run (bhs_fp64_peak); "
                           "MEASURED_PEAKS.json carries no FP64 figure",
            "traffic": traffic.get("traffic_bytes"), "traffic_note": traffic.get("launch", "no ncu capture found")
            + " -- dram__bytes_read.sum + dram__bytes_write.sum of that one launch (algorithmic bytes of the same launch: "
            + str(traffic.get("algorithmic_bytes")) + ")",
            "launches": int(g["count"]), "avg_launch_ms": g["ms"] / max(g["count"], 1),
            "flops_per_launch_avg": g["work"] / max(g["count"], 1),
            "scope": "trailing updates of the 128-wide outer blocks (97.7 % of the LU flops), every launch of the "
                     "profiled pass; the K = 32/64 updates inside the panel recursion are reported under inner_updates",
            "inner_updates": {"tflops": gi["work"] / (gi["ms"] * 1e-3) * 1e-12 if gi["ms"] > 0 else None,
                              "launches": int(gi["count"]), "ms_per_system": gi["ms"] / nprof},
        }
        # the same kernel on the largest trailing-update shape of a C3 factorisation, launched back to back (steady clocks;
        # the packing kernels of the stand-alone entry point are inside the bracket)
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
        out["roofline"]["isolated_largest_launch"] = {"shape": f"{Mt}x{Mt}x128", "tflops": iso_tf, "frac": iso_tf / peak_dmma}
        del La, Ua, Ca, wk
        out["kernel_split_ms_per_system"] = {n: v["ms"] / nprof for n, v in prof.items()}
        out["kernel_split_ms_per_system"]["eager_total"] = tot_ms / nprof
        out["lu"] = {"tflops": (8.0 / 3.0) * N ** 3 / (lu_ms / nprof * 1e-3) * 1e-12 if lu_ms > 0 else None,
                     "ms": lu_ms / nprof, "fp64_dmma_peak_tflops": peak_dmma, "fp64_dfma_peak_tflops": peak_dfma}
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
            nterms, _ = _ops.get_plan(3, N_END).coupling_stats()
            counted = (B * B - B) * (8.0 * nterms + 12.0 * H * H)  # 8 flop per coupling term + 12 per entry, per pair
            ctf = counted / (am["ms"] / nprof * 1e-3) * 1e-12
            out["assembly"] = {"gbs": gbs, "ms": am["ms"] / nprof, "hbm_peak_gbs": hbm, "frac_hbm": gbs / hbm,
                               "counted_tflops": ctf, "frac_fp64_counted": ctf / peak_dfma,
                               "counted_note": "SURVEY 8d flop count per ordered pair; the kernel executes it once per DISTINCT "
                                               "translation vector (48 of 240 pairs on this grid)",
                               "write_stream_gbs": wbest, "frac_write_stream": gbs / wbest,
                               "note": "hbm_peak_gbs is MEASURED_PEAKS.json's copy figure (read + write bytes); "
                                       "write_stream_gbs is a 4 GiB fill measured in this run"}
        del A, work, bufs

    # ---- u_scat points/s on the C5 field grid (64 spheres, n_end = 24), field rows split over the ranks -----
    if not args.no_uscat_grid:
        out_us = _bench_uscat(args, bhs, _ops, torch, dev, rank, world, barrier, dist if world > 1 else None)
        if rank == 0:
            out["uscat"] = out_us

    # ---- CPU baseline beside it (rank 0, N = 1 only) ------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import biem_oracle as O

        O.coupling_matrix("ba", N_END)
        ncpu = max(1, args.cpu_systems)
        sel = [ks_all[(i * 37) % len(ks_all)] for i in range(ncpu)]
        dt, _ = _oracle_systems(sel)
        out["cpu_baseline"] = {
            "value": ncpu / dt, "unit": "systems/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{ncpu} systems of the sweep (k = {', '.join(f'{v:.3f}' for v in sel)}), NumPy/SciPy oracle, "
                      f"LAPACK zgesv with all host cores",
        }
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


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


def _bench_uscat(args, bhs, _ops, torch, dev, rank, world, barrier, dist):
    """C5 field evaluation: 8x8 grid of 64 unit spheres, n_end = 24, k = 1, 2048^2 points on x2 = 0.

    The density is a deterministic synthetic vector with the decay of a solved one (solving the 36 864-unknown
    system is the serial part of C5 and is measured separately by tools/bench_c5.py); field rows are split in
    contiguous tiles over the ranks, exactly as the multi-GPU heat map does after the density broadcast."""
    from biem_helmholtz_sphere_b200.geometry import field_grid, grid_centers

    n_end, half, k = 24, 4, 1.0
    cen = torch.as_tensor(grid_centers(half, 3), device=dev)
    B = cen.shape[0]
    rad = torch.ones(B, dtype=torch.float64, device=dev)
    H = n_end * n_end
    rng = np.random.default_rng(0)
    deg = np.repeat(np.arange(n_end), 2 * np.arange(n_end) + 1)
    dens_np = (rng.standard_normal((B, H)) + 1j * rng.standard_normal((B, H))) * np.exp(-0.7 * deg)[None, :]
    dens = torch.as_tensor(dens_np, device=dev)
    if dist is not None:
        dist.broadcast(dens, src=0)  # the one collective of the path: solved coefficients -> every GPU
    G = args.grid
    rows = np.array_split(np.arange(G), world)[rank]
    x_np = field_grid(G, 20.0, 3)[:, rows, :].reshape(3, -1)
    x = torch.as_tensor(np.ascontiguousarray(x_np), device=dev)
    P = x.shape[1]
    work = _ops._work(_ops.load().bhs_uscat_workspace(_ops.get_plan(3, n_end).handle, B))
    for _ in range(3):
        o = _ops.uscat(3, n_end, cen, rad, k, 1.0, dens, x, work=work)
    reps = 5
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        o = _ops.uscat(3, n_end, cen, rad, k, 1.0, dens, x, work=work)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    # e2e: host grid in, host field out
    x_pin = torch.as_tensor(np.ascontiguousarray(x_np)).pin_memory()
    o_pin = torch.empty((P,), dtype=torch.complex128).pin_memory()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        xd = x_pin.to(dev, non_blocking=True)
        od = _ops.uscat(3, n_end, cen, rad, k, 1.0, dens, xd, work=work)
        o_pin.copy_(od, non_blocking=True)
        torch.cuda.synchronize()
    te = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    nan_frac = float(torch.isnan(o.real).double().mean())
    peak_dfma = max(_ops.fp64_peak(0, 4096), 1e-9) if rank == 0 else 1.0
    Ptot = G * G
    flops = 8.0 * Ptot * B * H
    return {
        "metric": "uscat points/sec (second half of BASELINE.json's metric)", "value": Ptot / (ms * 1e-3), "unit": "points/s", "ms": ms,
        "workload": f"C5: 64 unit spheres (8x8 grid), n_end=24 (H=576), k=1, {G}x{G} field grid on x2=0 over [-20,20]^2, "
                    f"rows split over {world} rank(s); synthetic density",
        "e2e": {"value": Ptot / (float(te[0]) * 1e-3), "unit": "points/s", "h2d_bytes": int(24 * Ptot), "d2h_bytes": int(16 * Ptot)},
        "roofline": {"bound": "fp64", "achieved": flops / (ms * 1e-3) * 1e-12 / world, "peak": peak_dfma, "unit": "TFLOP/s",
                     "frac": flops / (ms * 1e-3) * 1e-12 / world / peak_dfma,
                     "note": "counted flops = 8 P B H (one complex FMA per point x ball x harmonic; special-function "
                             "generation not counted); per-GPU figure against the DFMA peak measured in this run. The grid and "
                             "the sphere centres are coplanar (x2 = 0), so the device-selected planar variant of the kernel runs "
                             "(8 FP64 instructions per step instead of 12); non-coplanar inputs measure 0.44"},
        "nan_fraction": nan_frac,
        "ncu": _load_profile("r01_uscat_ncu.json"),
    }


if __name__ == "__main__":
    a = _args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
