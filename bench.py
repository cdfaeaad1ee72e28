#!/usr/bin/env python
"""
bench.py — DR-CVaR safe-halfspace throughput on B200 (BASELINE.json metric).

  python bench.py --gpus 1 --steps K --warmup W                       (this repo's CUDA path)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   (scenario-sharded, weak scaling)
  python bench.py --impl reference ...                                 (CPU arm: oracle port on all host cores)

A "step" is one pass of the hot path over the whole resident synthetic batch (BASELINE config 4:
4096 scenarios x 8 obstacles x horizon 20 = 655 360 halfspaces x N = 10 000 samples, alpha 0.1, eps 0.01),
one kernel launch per step per GPU.  Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DR-CVaR halfspaces/sec at N=10k samples"
UNIT = "halfspaces/s"
RISK = dict(alpha=0.1, delta=0.1, epsilon=0.01, robot_radius=0.3, obstacle_radius=0.3)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--scenarios", type=int, default=4096)
    ap.add_argument("--obstacles", type=int, default=8)
    ap.add_argument("--horizon", type=int, default=20)
    ap.add_argument("--samples", type=int, default=10000)
    ap.add_argument("--dtype", choices=["f32", "f64"], default="f32", help="sample input dtype (arithmetic is fp64)")
    ap.add_argument("--e2e-halfspaces", type=int, default=4096, help="host batch per end-to-end step")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="budget of the cpu_baseline leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-generated", action="store_true", help="skip the generate-mode leg (samples drawn in-kernel)")
    ap.add_argument("--no-large-n", action="store_true", help="skip the N = 100 000 leg (cluster / DSMEM kernel)")
    ap.add_argument("--large-n-halfspaces", type=int, default=16384, help="halfspaces of the N = 100 000 leg")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic batch: {a.scenarios} scenarios x {a.obstacles} obstacles x horizon {a.horizon} x "
            f"N={a.samples} samples ({a.dtype} inputs), alpha={RISK['alpha']}, eps={RISK['epsilon']}")


def algorithmic_bytes_per_halfspace(n, elem):
    # one read of the samples + ego in + h[2], g[3] out
    return n * 2 * elem + 16 + 40


# ------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler:
    """Samples SM clock / throttle reasons DURING the timed region (NVML, 20 ms period)."""

    def __init__(self, torch_device_index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            import torch
            self.nv = pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(torch_device_index).uuid)
                if not uuid.startswith("GPU-"):
                    uuid = "GPU-" + uuid
                h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = torch_device_index
                if vis:
                    try:
                        idx = int(vis.split(",")[torch_device_index])
                    except Exception:
                        pass
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._h = h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                r = int(get_reasons(self._h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=2)
        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "n_samples": len(self.samples),
        }


# ------------------------------------------------------------------------------------------ synthetic input
def synth_means(B, seed):
    """mu_b uniform in the 1-5 m annulus around the ego (origin)."""
    import numpy as np
    rng = np.random.RandomState(seed)
    r = 1.0 + 4.0 * rng.rand(B)
    th = 2.0 * math.pi * rng.rand(B)
    return np.stack([r * np.cos(th), r * np.sin(th)], axis=1)


def make_device_batch(B, N, dtype, device, seed, chunk=2048):
    """samples[b] = mu_b + 0.1 z, z ~ N(0, I2)  (noise cov diag(0.01, 0.01): simulation/obstacles.py:134)."""
    import torch
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    mu = torch.from_numpy(synth_means(B, seed)).to(device)
    s = torch.empty((B, N, 2), dtype=dtype, device=device)
    for b0 in range(0, B, chunk):
        nb = min(chunk, B - b0)
        z = torch.randn((nb, N, 2), generator=gen, dtype=torch.float32, device=device)
        s[b0:b0 + nb] = (mu[b0:b0 + nb, None, :] + 0.1 * z.double()).to(dtype) if dtype == torch.float64 else \
            (mu[b0:b0 + nb, None, :].float() + 0.1 * z)
        del z
    ego = torch.zeros((B, 2), dtype=torch.float64, device=device)
    return s, ego


# ------------------------------------------------------------------------------------------ CPU legs (oracle)
_CPU_DATA = {}


def _cpu_chunk(args):
    lo, hi = args
    from oracle import closed_form as cf
    s, ego = _CPU_DATA["s"], _CPU_DATA["ego"]
    out = []
    for b in range(lo, hi):
        o = cf.halfspace(s[b], ego[b], RISK["alpha"], RISK["delta"], RISK["epsilon"], RISK["robot_radius"],
                         RISK["obstacle_radius"])
        out.append((o.g_mean, o.g_cvar, o.g_dr))
    return out


def cpu_baseline_leg(samples_np, ego_np, seconds):
    """Oracle (numpy closed-form port of the reference path) on ONE core: cycles over the sample for `seconds`."""
    from oracle import closed_form as cf
    nb = samples_np.shape[0]
    t0 = time.perf_counter()
    res, n = [], 0
    while True:
        b = n % nb
        o = cf.halfspace(samples_np[b], ego_np[b], RISK["alpha"], RISK["delta"], RISK["epsilon"],
                         RISK["robot_radius"], RISK["obstacle_radius"])
        if n < nb:
            res.append((o.g_mean, o.g_cvar, o.g_dr))
        n += 1
        if time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return n / dt, n, dt, res


def run_reference_arm(a):
    """--impl reference: the CPU restatement of the reference path on all host cores (oracle port)."""
    import multiprocessing as mp
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = len(os.sched_getaffinity(0))
    per_step = 64 * cores
    N = a.samples
    rng = np.random.RandomState(42)
    mu = synth_means(per_step, 42)
    s = mu[:, None, :] + 0.1 * rng.standard_normal((per_step, N, 2))
    s = s.astype(np.float32 if a.dtype == "f32" else np.float64)
    _CPU_DATA["s"] = s
    _CPU_DATA["ego"] = np.zeros((per_step, 2))
    chunks = [(i * 64, (i + 1) * 64) for i in range(cores)]
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        for _ in range(max(a.warmup, 1)):
            pool.map(_cpu_chunk, chunks)
        t0 = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_cpu_chunk, chunks)
        dt = time.perf_counter() - t0
    value = per_step * a.steps / dt
    sample = (f"{per_step} halfspaces per step (64 per core) of the same workload, numpy closed-form port of the "
              f"reference's LP path (the reference's cvxpy/ECOS solver is not installable here)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(a), "halfspaces_per_step": per_step, "samples_per_halfspace": N},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 as pkg
    from dr_cvar_mpc_safety_filter_motion_planning_collison_avoidance_b200 import _lib, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    tdtype = torch.float32 if a.dtype == "f32" else torch.float64
    elem = 4 if a.dtype == "f32" else 8
    N = a.samples
    per_scn = sharding.halfspaces_per_scenario(a.obstacles, a.horizon)
    # weak scaling: every GPU owns a full config-4-sized shard (scenarios are independent; no collective on the path)
    B = a.scenarios * per_scn
    need = B * N * 2 * elem
    free, _total = torch.cuda.mem_get_info(device)
    if need > free - (8 << 30):
        scn = max(1, int((free - (8 << 30)) // (per_scn * N * 2 * elem)))
        a.scenarios = scn
        B = scn * per_scn
    samples, ego = make_device_batch(B, N, tdtype, device, seed=42 + rank)
    out = None
    stream = torch.cuda.current_stream(device)

    def step():
        nonlocal out
        out = pkg.compute_halfspaces(samples, ego, stream=stream, out=out, **RISK)

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)

    sampler = ClockSampler(local_rank)
    launches0 = pkg.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.start()
    start.record(stream)
    for k in range(a.steps):
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
    stop.record(stream)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    clocks = sampler.stop()
    launches = pkg.launch_count() - launches0
    total_ms = start.elapsed_time(stop)
    kern_ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    t = torch.tensor([total_ms, statistics.mean(kern_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kern_ms_avg = float(t[0]), float(t[1])
    value = world * B * a.steps / (total_ms * 1e-3)

    # ---- final gather of the halfspaces (the only exchange; not on the hot path)
    gather_ms = 0.0
    if world > 1:
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        gh, gg = sharding.gather_results(out.h, out.g, a.scenarios * world, per_scn)
        torch.cuda.synchronize(device)
        gather_ms = (time.perf_counter() - t0) * 1e3
        assert gh.shape[0] == world * B

    # ---- end-to-end through the public API with HOST (pinned) buffers
    Be = min(a.e2e_halfspaces, B)
    host = torch.empty((Be, N, 2), dtype=tdtype, pin_memory=True)
    host.copy_(samples[:Be])
    ego_h = np.zeros((Be, 2))
    host_np = host.numpy()
    e2e_steps = max(3, min(a.steps, 10))
    for _ in range(2):
        hres = pkg.compute_halfspaces(host_np, ego_h, **RISK)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        hres = pkg.compute_halfspaces(host_np, ego_h, **RISK)
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * Be * e2e_steps / float(te[0])
    h2d = Be * N * 2 * elem + Be * 16
    d2h = Be * (16 + 16 + 24 + 8 + 8 + 8 + 4)
    torch.cuda.synchronize(device)
    assert np.array_equal(hres.g, out.g[:Be].cpu().numpy()), "host path and device path disagree"

    # ---- generate mode: the same batch with the samples drawn inside the kernel — no sample bytes
    #      in HBM or over PCIe; compute-bound (Philox + Box-Muller), reported beside the headline, not instead of it
    generated = None
    if a.dtype == "f32" and not a.no_generated and N <= pkg.max_samples(np.float32):
        import ctypes as C
        lib = _lib.load()
        g = torch.Generator(device=device).manual_seed(4242 + rank)
        ang = torch.rand(B, generator=g, device=device, dtype=torch.float64) * (2 * np.pi)
        rad = 1.0 + 4.0 * torch.rand(B, generator=g, device=device, dtype=torch.float64)
        mean_t = torch.stack([rad * torch.cos(ang), rad * torch.sin(ang)], dim=1).contiguous()
        chol_t = torch.tensor([0.1, 0.0, 0.1], dtype=torch.float64, device=device).repeat(B, 1).contiguous()
        gout = pkg.HalfspaceBatch(h=torch.empty_like(out.h), h_mean=torch.empty_like(out.h_mean), g=torch.empty_like(out.g),
                                  cvar=None, var=None, g_star=None, status=torch.zeros_like(out.status))

        def gen_step():
            rc = lib.drcvar_halfspaces_generated_f32(
                mean_t.data_ptr(), chol_t.data_ptr(), 42, rank * B, B, N, ego.data_ptr(), None,
                RISK["alpha"], RISK["delta"], RISK["epsilon"], RISK["robot_radius"], RISK["obstacle_radius"], 0,
                gout.h.data_ptr(), gout.h_mean.data_ptr(), gout.g.data_ptr(), None, None, None,
                gout.status.data_ptr(), None, None, local_rank, C.c_void_p(stream.cuda_stream))
            _lib.check(rc)

        gen_step()
        torch.cuda.synchronize(device)
        g_steps = 3
        gs, ge = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gs.record(stream)
        for _ in range(g_steps):
            gen_step()
        ge.record(stream)
        torch.cuda.synchronize(device)
        g_ms = gs.elapsed_time(ge) / g_steps
        # end to end from HOST inputs (nominal positions, covariance, ego: 56 B per halfspace) to HOST outputs
        Bg = min(B, 65536)
        mean_h, ego_g = mean_t[:Bg].cpu().numpy(), np.zeros((Bg, 2))
        cov_h = np.diag([0.01, 0.01])
        pkg.compute_halfspaces_generated(mean_h, cov_h, N, 42, ego=ego_g, **RISK)
        t0 = time.perf_counter()
        for _ in range(3):
            gres = pkg.compute_halfspaces_generated(mean_h, cov_h, N, 42, ego=ego_g, **RISK)
        g_e2e = 3 * Bg / (time.perf_counter() - t0)
        generated = {"value": world * B / (g_ms * 1e-3), "unit": UNIT, "ms_per_launch": g_ms,
                     "samples_per_s": world * B * N / (g_ms * 1e-3),
                     "e2e": {"value": world * g_e2e, "unit": UNIT, "halfspaces_per_step": Bg,
                             "h2d_bytes_per_step": Bg * 56, "d2h_bytes_per_step": Bg * 84},
                     "fallback_halfspaces": int((gres.status & 2 != 0).sum()),
                     "note": "samples drawn in-kernel (Philox4x32-10 + fp32 Box-Muller, oracle/sample_gen.py); "
                             "per-rank numbers scaled by world size"}

    # ---- BASELINE config 5 sample count (N = 100 000) on a shard that fits next to the headline batch: the cluster /
    #      DSMEM kernel (one cluster of 4 CTAs per halfspace, every sample read once) beside the two-pass streaming kernel
    large_n = None
    if a.dtype == "f32" and not a.no_large_n and N <= 32768:
        NL, BL = 100000, a.large_n_halfspaces
        free, _t = torch.cuda.mem_get_info(device)
        BL = int(max(0, min(BL, (free - (6 << 30)) // (NL * 8))))
        if BL >= 1024:
            sl, egl = make_device_batch(BL, NL, torch.float32, device, seed=777 + rank, chunk=256)
            bytes_l = algorithmic_bytes_per_halfspace(NL, 4) * BL
            large_n = {"samples_per_halfspace": NL, "halfspaces_per_gpu": BL, "unit": UNIT}
            for name, fl in (("cluster_kernel", 0), ("streaming_kernel", _lib.FLAG_NO_CLUSTER)):
                lo = None
                for _ in range(2):
                    lo = pkg.compute_halfspaces(sl, egl, stream=stream, out=lo, flags=fl, **RISK)
                torch.cuda.synchronize(device)
                ls, le = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                ls.record(stream)
                for _ in range(3):
                    lo = pkg.compute_halfspaces(sl, egl, stream=stream, out=lo, flags=fl, **RISK)
                le.record(stream)
                torch.cuda.synchronize(device)
                l_ms = ls.elapsed_time(le) / 3
                large_n[name] = {"value": world * BL / (l_ms * 1e-3), "ms_per_launch": l_ms,
                                 "hbm_gbs": bytes_l / (l_ms * 1e-3) / 1e9,
                                 "general_path_halfspaces": int((lo.status & 2 != 0).sum().item())}
                if name == "cluster_kernel":
                    keep = (lo.h.clone(), lo.var.clone())
                else:
                    large_n["bit_identical_h_and_T"] = bool(torch.equal(keep[0], lo.h) and torch.equal(keep[1], lo.var))
            del sl, egl

    # ---- roofline of the (single) kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    alg_bytes = algorithmic_bytes_per_halfspace(N, elem) * B
    achieved = alg_bytes / (kern_ms_avg * 1e-3) / 1e9
    traffic = None
    for tname in ("traffic.json", "traffic_n100k.json"):   # DRAM bytes per halfspace from the committed ncu captures
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if tj.get("samples") == N and tj.get("dtype") == a.dtype:
                    traffic = tj["dram_bytes_per_halfspace"] * B
            except Exception:
                pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "halfspace_kernel" if N <= pkg.max_samples(np.float32 if a.dtype == "f32" else np.float64)
                else ("cluster_kernel_f32" if a.dtype == "f32" and N > 32768 else "streaming_kernel"),
                "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms_avg": kern_ms_avg}

    # ---- CPU baseline beside it (rank 0, N = 1 GPU only) + parity spot-check of the timed results
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        nb = min(B, 1024)
        s_np = samples[:nb].cpu().numpy()
        e_np = ego[:nb].cpu().numpy()
        v, n_done, dt, res = cpu_baseline_leg(s_np, e_np, a.cpu_seconds)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"{n_done} oracle evaluations cycling over the first {nb} halfspaces of the same batch in "
                         f"{dt:.1f} s, numpy closed-form port of the reference path (1 core)"}
        g_gpu = out.g[:len(res)].cpu().numpy()
        ref = np.array(res)
        parity = {"halfspaces": len(res), "tolerance": "1e-6 m (fp32 inputs) / 1e-9 rel (fp64 inputs)",
                  "max_rel_err": float(np.max(np.abs(g_gpu - ref) / np.maximum(1.0, np.abs(ref))))}

    if rank == 0:
        line = {
            "metric": METRIC if N == 10000 else f"DR-CVaR halfspaces/sec at N={N} samples", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": total_ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "halfspaces_per_gpu": B, "samples_per_halfspace": N,
                       "input_dtype": a.dtype, "parallelism": f"scenario-shard x{world} (no collective on the hot path)",
                       "l2": f"inputs ({need / 1e9:.1f} GB per GPU) are larger than L2; no flush needed"},
            "roofline": roofline, "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "halfspaces_per_step": Be, "steps": e2e_steps,
                    "note": "compute_halfspaces() on pinned host numpy buffers: chunked H2D + kernel + D2H inside the timed region"},
            "gpu_launches": launches, "clocks": clocks, "hbm_gbs_aggregate": achieved * world, "generated": generated,
            "large_n": None if large_n is None else dict(large_n, hbm_peak_gbs=peak, note="per-rank shard, scaled by world size"),
            "gather_ms": gather_ms, "parity_spot_check": parity,
            "status_fallback_halfspaces": int((out.status != 0).sum().item()),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference_arm(a)
    return run_ours(a)


if __name__ == "__main__":
    sys.exit(main())
