#!/usr/bin/env python
"""bench.py — headline benchmark of the GP-surrogate + acquisition hot path.

metric (BASELINE.json): candidate posterior+EI evaluations / second at n = 8192 observations,
d = 20 (config C4 of SURVEY §8d: SE kernel, l = 1, sigma^2 = 1, noise 1e-2, 2,097,152 candidates
per GPU = 16 M over 8 GPUs).  A "step" is one fused sweep (K* tiles -> mean, L^-1 K* on the FP64
tensor pipe -> variance -> EI -> stable top-100) over the rank's candidate shard.

  value : candidates/s, whole job (all ranks), candidates already resident in HBM
  e2e   : the same through the host-buffer C-ABI call abo_acq_eval (pinned host candidates,
          H2D copy, sweep, D2H of the scores, top-k) — copies inside the timed region
  roofline : the dominant kernel (DMMA triangular product + sum of squares), algorithmic
          n^2 flop per candidate / its CUDA-event duration, against the measured FP64 GEMM peak
  cholesky : second BASELINE metric — blocked FP64 Cholesky TFLOP/s at n = 8192
  cpu_baseline : the CPU restatement (oracle, NumPy/SciPy on OpenBLAS, all host cores) on a
          bounded sample of the same workload (the Julia reference cannot run here: no Julia)

`--impl reference` times that CPU restatement as the reference arm.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_OBS, DIM = 8192, 20
M_PER_GPU = 2_097_152
TOPK = 100
METRIC = "candidate posterior+EI evals/sec at n=8192,d=20"
UNIT = "candidates/s"


def flops_per_candidate(n, d):
    return float(n) * n + 3.0 * n * d + 4.0 * n


def fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) has no FP64 row, so the
    denominator is this repo's own measurement on the same pool: cuBLAS Dgemm NT 8192^3
    (tools/fp64_peaks.cu -> profiles/fp64_peaks_r01.json)."""
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if "fp64_tflops" in mp:
            return float(mp["fp64_tflops"]), "MEASURED_PEAKS.json fp64_tflops"
    except Exception:
        pass
    try:
        pk = json.load(open(os.path.join(ROOT, "profiles", "fp64_peaks_r01.json")))
        return float(pk["cublas_dgemm_nt_tflops"]["8192"]), \
            "cuBLAS Dgemm NT 8192^3 measured on this pool (profiles/fp64_peaks_r01.json); MEASURED_PEAKS.json has no FP64 row"
    except Exception:
        return 37.0, "nominal B200 FP64 (no measurement file found)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1; the CPU legs are meant to use every host core."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count() or 1)
    except Exception:
        pass


def cpu_restatement_throughput(sample_m, seed=42):
    """Oracle (CPU restatement of the reference path) on a bounded sample of the C4 workload:
    conditioning is untimed set-up, the timed part is posterior mean + variance + EI over
    `sample_m` candidates (best of 2 after one warm-up), all host cores."""
    from oracle import abo_oracle as orc
    use_all_host_cores()
    c = orc.make_config("C4", seed=seed, n=N_OBS, m=sample_m, d=DIM)
    t0 = time.perf_counter()
    post = orc.fit_standard(c["X"], c["y"], c["kind"], c["inv_ls"], c["scale"], c["noise"])
    t_fit = time.perf_counter() - t0
    best = float(c["y"].min())
    times = []
    for it in range(3):
        t0 = time.perf_counter()
        mu, var = orc.posterior_mean_var(post, c["Xc"], chunk=2048)
        ei = orc.expected_improvement(mu, var, 0.01, best)
        idx = orc.sortperm_rev(ei, TOPK)
        times.append(time.perf_counter() - t0)
    t = min(times[1:])
    return sample_m / t, t, t_fit


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path.  The reference is Julia
    and cannot run in this image (no Julia toolchain, no network); the timed stand-in is the CPU
    restatement in oracle/ (NumPy/SciPy on OpenBLAS with all host cores), kind = "port"."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.cpu_sample
    from oracle import abo_oracle as orc
    use_all_host_cores()
    c = orc.make_config("C4", n=N_OBS, m=sample, d=DIM)
    post = orc.fit_standard(c["X"], c["y"], c["kind"], c["inv_ls"], c["scale"], c["noise"])
    best = float(c["y"].min())

    def step():
        mu, var = orc.posterior_mean_var(post, c["Xc"], chunk=2048)
        ei = orc.expected_improvement(mu, var, 0.01, best)
        orc.sortperm_rev(ei, TOPK)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"C4: StandardGP SE n={N_OBS} d={DIM}, EI + top-{TOPK}; each step a bounded sample of "
                               f"{sample} candidates on the host cores", "n": N_OBS, "d": DIM},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} candidates per step, oracle/abo_oracle.py (NumPy/SciPy OpenBLAS); "
                                   "the Julia reference cannot run in this image"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cands", dest="m", type=int, default=M_PER_GPU, help="candidates per GPU per step")
    ap.add_argument("--n", type=int, default=N_OBS)
    ap.add_argument("--cpu-sample", type=int, default=8192)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cholesky", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import abo_b200 as abo
    from oracle import abo_oracle as orc            # data generators + cpu_baseline leg only

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = abo.default_context(local_rank)
    n, d, m = args.n, DIM, args.m

    # ---- conditioning set: identical on every rank (seeded); fitted redundantly per rank —
    # deterministic kernels make the replicas bit-identical — un-timed set-up.
    c = orc.make_config("C4", n=n, m=1, d=d)
    model = abo.StandardGP(c["scale"] * abo.with_lengthscale(abo.SqExponentialKernel(), 1.0 / c["inv_ls"]), c["noise"],
                           ctx=ctx)
    t0 = time.perf_counter()
    t_sync = None
    if world == 1:
        model = abo.update(model, c["X"], c["y"])
    else:
        # rank 0 conditions the surrogate; L, L^-1, X, alpha and the hyper-parameters reach the other
        # ranks with one NCCL broadcast over NVLink (abo_gp_sync) — set-up, outside the timed region
        abo.init_nccl_context(ctx)
        model = abo.update(model, c["X"], c["y"]) if rank == 0 else abo.empty_posterior_like(model, d)
        torch.cuda.synchronize(); dist.barrier()
        abo.sync_posterior(model, 0)                 # first call also sets the NCCL channels up
        torch.cuda.synchronize(); dist.barrier()
        ts = time.perf_counter()
        abo.sync_posterior(model, 0)
        torch.cuda.synchronize(); dist.barrier()
        t_sync = time.perf_counter() - ts
    t_fit = time.perf_counter() - t0
    acq = abo.ExpectedImprovement(0.01, float(c["y"].min()))
    params = acq.params()
    h = model.gpx

    # ---- candidates: this rank's shard, device-resident for `value`, pinned host copy for e2e
    gen = torch.Generator(device="cuda")
    gen.manual_seed(1234 + rank)
    Xc_dev = torch.rand((m, d), dtype=torch.float64, device="cuda", generator=gen)
    scores_dev = torch.empty(m, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    lib_stream = torch.cuda.ExternalStream(ctx.stream())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev():
        return h.acq_eval_dev(acq.acq_id, params, Xc_dev.data_ptr(), m, scores_dev.data_ptr(), k=TOPK)

    for _ in range(args.warmup):
        step_dev()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    for _ in range(args.steps):
        top_idx, top_val = step_dev()
    e1.record(lib_stream)
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count() - launches0
    # one extra, un-timed pass with per-kernel CUDA-event brackets (serialises the K* builder, which
    # otherwise overlaps the contraction) for the roofline of the dominant kernel
    ctx.profile(True)
    step_dev()
    prof_ms, prof_n = ctx.profile_read()
    ctx.profile(False)
    prof_steps = 1
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * m * args.steps / (ms_max * 1e-3)

    # ---- e2e: host buffers through the C-ABI call, copies inside the timed region
    Xc_host = torch.empty((m, d), dtype=torch.float64, pin_memory=True)
    Xc_host.copy_(Xc_dev)
    Xh = Xc_host.numpy()
    scores_host = torch.empty(m, dtype=torch.float64, pin_memory=True)
    import ctypes as C
    from abo_b200 import _lib
    ti = np.empty(TOPK, dtype=np.int64); tv = np.empty(TOPK)
    pp = _lib.f64(params)

    def step_e2e():
        _lib.check(_lib.lib().abo_acq_eval(h._h, acq.acq_id, _lib.ptr(pp), C.c_void_p(Xh.ctypes.data), m,
                                           C.c_void_p(scores_host.data_ptr()), TOPK, _lib.ptr(ti), _lib.ptr(tv)))
    step_e2e()
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record(lib_stream)
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    e3.record(lib_stream)
    barrier()
    wall_e2e = time.perf_counter() - w0
    ms_e2e = max(e2.elapsed_time(e3), 1e3 * wall_e2e)          # host-side top-k is part of the call
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * m * args.steps / (float(t.item()) * 1e-3)
    same_top = bool(np.array_equal(ti, top_idx))

    # ---- roofline of the dominant kernel (rank 0's numbers)
    peak, peak_src = fp64_peak()
    Npad = (n + 127) // 128 * 128
    trmm_ms, trmm_n = prof_ms[1], max(prof_n[1], 1)
    cand_per_launch = m * prof_steps / trmm_n
    achieved = (float(n) * n * cand_per_launch) / (trmm_ms / trmm_n * 1e-3) / 1e12
    traffic = None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_sweep_tma_r01_summary.json")))[0]
        if n == N_OBS and abs(cand_per_launch - 4096) < 1:      # the capture was taken on this exact launch shape
            def _b(x):
                v, u = x.split()[:2]
                return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            traffic = _b(summ["dram__bytes_read.sum"]) + _b(summ["dram__bytes_write.sum"])
    except Exception:
        traffic = None
    roofline = {"bound": "tensor", "kernel": "sweep_tma_kernel (TMA + mbarrier + DMMA: W = L^-1 K*, fused column sum of squares)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_flop_per_candidate": float(n) * n,
                "candidates_per_launch": cand_per_launch, "avg_launch_ms": trmm_ms / trmm_n,
                "share_of_step": {"ks_build": prof_ms[0] / max(sum(prof_ms), 1e-9),
                                  "trmm_sumsq": prof_ms[1] / max(sum(prof_ms), 1e-9),
                                  "acq_epilogue": prof_ms[2] / max(sum(prof_ms), 1e-9)},
                "whole_step_tflops": flops_per_candidate(n, d) * m * args.steps / (ms * 1e-3) / 1e12}

    # ---- Cholesky TFLOP/s (second BASELINE metric), rank 0 only
    chol = None
    if rank == 0 and not args.no_cholesky:
        Xd = torch.from_numpy(c["X"]).cuda()
        K0 = torch.exp(-0.5 * torch.cdist(Xd, Xd) ** 2) + 1e-2 * torch.eye(n, dtype=torch.float64, device="cuda")
        A = torch.empty_like(K0)
        best = 1e30
        for it in range(4):
            A.copy_(K0)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(lib_stream)
            ctx.potrf_dev(A.data_ptr(), Npad if Npad == n else n, n)
            s1.record(lib_stream)
            torch.cuda.synchronize()
            if it:
                best = min(best, s0.elapsed_time(s1))
        fl = n ** 3 / 3.0 + n ** 2 / 2.0
        chol = {"n": n, "ms": best, "tflops": fl / (best * 1e-3) / 1e12, "frac_of_peak": fl / (best * 1e-3) / 1e12 / peak,
                "first_fit_s_incl_allocation": t_fit}
        del K0, A, Xd
        # the whole conditioning step (K build + Cholesky + triangular inverse + alpha), warm, through the host API
        tw = []
        for _ in range(3):
            t1 = time.perf_counter(); abo.update(model, c["X"], c["y"], allow_append=False); tw.append(time.perf_counter() - t1)
        chol["fit_ms_warm"] = 1e3 * min(tw)
        # a larger factorisation: the look-ahead schedule approaches the GEMM rate as the panel share shrinks
        n2 = 2 * n
        X2 = torch.rand((n2, DIM), dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(7))
        K2 = torch.exp(-0.5 * torch.cdist(X2, X2) ** 2) + 1e-2 * torch.eye(n2, dtype=torch.float64, device="cuda")
        A2 = torch.empty_like(K2)
        best2 = 1e30
        for it in range(3):
            A2.copy_(K2)
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record(lib_stream)
            ctx.potrf_dev(A2.data_ptr(), n2, n2)
            s1.record(lib_stream)
            torch.cuda.synchronize()
            if it:
                best2 = min(best2, s0.elapsed_time(s1))
        fl2 = n2 ** 3 / 3.0 + n2 ** 2 / 2.0
        chol["larger"] = {"n": n2, "ms": best2, "tflops": fl2 / (best2 * 1e-3) / 1e12,
                          "frac_of_peak": fl2 / (best2 * 1e-3) / 1e12 / peak}
        del K2, A2, X2

    cpu = None
    if rank == 0 and args.gpus == 1 and not args.no_cpu_baseline:
        v, tstep, tfit = cpu_restatement_throughput(args.cpu_sample)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": f"{args.cpu_sample} candidates (posterior mean+var+EI+top-{TOPK}) at n={n}, d={d}; "
                         f"{tstep:.2f} s per pass; oracle/abo_oracle.py on OpenBLAS (Julia reference cannot run here)",
               "fit_s": tfit}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"C4: StandardGP SE n={n} d={d} noise=1e-2, EI(xi=0.01) + stable top-{TOPK} over "
                                   f"{m} candidates per GPU ({world * m} total); candidate set ({m * d * 8 / 1e6:.0f} MB) "
                                   f"and K* tiles exceed L2, no explicit flush",
                       "n": n, "d": d, "candidates_per_gpu": m, "l2": "inputs larger than L2"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": m * d * 8, "d2h_bytes_per_step": m * 8,
                    "topk_matches_device_run": same_top},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "posterior_broadcast": None if t_sync is None else {
                "bytes": 2 * Npad * Npad * 8, "ms": 1e3 * t_sync, "gb_per_s": 2 * Npad * Npad * 8 / t_sync / 1e9,
                "how": "abo_gp_sync: NCCL broadcast of L and L^-1 (+ X, alpha) from rank 0, second call"},
            "cholesky": chol,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    del model                                      # release the device state explicitly, not at interpreter exit
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
