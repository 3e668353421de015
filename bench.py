#!/usr/bin/env python
"""bench.py — log-posterior evals/s of the epi-mcmc hot path on N B200s.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path (calclogl(transformParams(theta)) + calclogp(theta))
over one batch of `--chains` proposals PER GPU (weak scaling; chains are independent, no
data-path collective).  Workload = BASELINE.json configs[2]: the age-structured model
(model-age-groups2q) on the synthetic A300 series, 65,536 chains per GPU, RK4 m = 4.

JSON line keys beyond the base contract:
  roofline      dominant kernel (pass-2 ODE) against the FP64 DFMA peak measured live
  cpu_baseline  the CPU oracle ("port": R is not installed) on a bounded sample, all host threads
  e2e           the same metric through the host-buffer C ABI (epi_eval_submit/wait, and epi_eval beside it)
                with pinned HOST buffers (H2D + D2H inside)
  kernels       mean device ms of the four launches of one step
  f_alg         algorithmic FLOP per evaluation (SURVEY.md §8d formula) and the whole-step fraction
  sampler       fused propose/accept DE-MC run: evals/s inside the sampler and ESS/s
"""
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

# DRAM bytes per launch of each kernel (read + write) from the committed ncu --set full capture of
# this exact workload (profiles/r1_v15_ncu_full.txt); only reported when the run matches that workload
NCU_TRAFFIC = {"ode_pass1": 296.4e6, "profiles_threshold": 301.0e6, "ode_pass2": 467.4e6, "score": 617.1e6}
# sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active of the same capture
NCU_FP64_PIPE_PCT = {"ode_pass1": 50.6, "profiles_threshold": 62.5, "ode_pass2": 47.9, "score": 29.9}
# FP64 flop actually EXECUTED per launch in the same capture: thread-level predicated-on DFMA x 2 + DMUL + DADD from the
# report's SASS table.  SURVEY 8d counts the reference's arithmetic (bounds-guard comparisons, the output variables
# Re/Rt, the removed compartment, per-call segment selection ...): the ODE kernels execute about a third of that, which
# is why their algorithmic `frac` exceeds 1 — the pipe utilisation above and this figure are the honest bound for them
NCU_EXECUTED_FLOP = {"ode_pass1": 2.749e9, "profiles_threshold": 4.985e9, "ode_pass2": 12.512e9, "score": 11.311e9}
NCU_WORKLOAD = ("A300", 65536, 4)

METRIC = "log_posterior_evals_per_s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A300")
    ap.add_argument("--chains", type=int, default=65536)
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--no-sampler", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sampler-iters", type=int, default=3000)
    ap.add_argument("--sampler-burnin", type=int, default=1000)
    ap.add_argument("--sampler-fulldim", action="store_true", help="DE-MC without subspace (crossover) updates")
    return ap.parse_args()


def make_theta(ds_name, n, seed, mn, mx, init):
    """proposals where a running sampler spends its time: +-0.5 % of the box width around the
    model file's init (= the mode of the synthetic data), so every chain does the full work
    (valid data_offset, pass 2 integrated).  Box-uniform draws would let ~1/3 of the chains skip
    pass 2 and flatter the number."""
    rng = np.random.default_rng(seed)
    th = init + (rng.uniform(size=(n, len(init))) - 0.5) * 0.01 * (mx - mn)
    return np.clip(th, mn, mx)


def f_alg(flavour, P, m, taps_mean, n_terms, n_prior):
    """ALGORITHMIC FLOP per evaluation, SURVEY.md §8d"""
    age = flavour in ("age2q", "age2i")
    S, K, C1, C2 = (10, 6, 2, 6) if age else (4, 2, 1, 2)
    P1 = 60 if flavour == "age2q" else P  # pass 1: 60 days in 2q, the whole period in 2i / simple
    F_rhs = 124 if age else 72
    steps = m * (P1 + P)
    ode = steps * (4 * F_rhs + 14 * S)
    total = ode + K * (taps_mean + 2) * 240 + (C2 * P + C1 * P1) * 2 * taps_mean + 6 * C2 * P + 72 * n_terms + 40 * n_prior
    per_kernel = {
        "ode_pass1": m * P1 * (4 * F_rhs + 14 * S),
        "profiles_threshold": K * (taps_mean + 2) * 240 + C1 * P1 * 2 * taps_mean,
        "ode_pass2": m * P * (4 * F_rhs + 14 * S),
        "score": C2 * P * 2 * taps_mean + 6 * C2 * P + 72 * n_terms + 40 * n_prior,
    }
    # the removed compartment(s) are not integrated on the hot path (bit-identical): the RK4
    # update of 2 (age) / 1 (simple) of the S states is not executed
    n_r = 2 if age else 1
    return dict(total=total, ode=ode, per_kernel=per_kernel, score_per_day=C2 * 2 * taps_mean + 6 * C2,
                ode_executed_share=(4 * F_rhs + 14 * (S - n_r)) / (4 * F_rhs + 14 * S))


def mean_taps(flavour, theta):
    if flavour not in ("age2q", "age2i"):
        return 16.0
    t = []
    if flavour == "age2q":
        prof = ((theta[:, 43], 5.0), (theta[:, 11], theta[:, 19]), (theta[:, 12], theta[:, 20]),
                (theta[:, 44], 5.0), (theta[:, 15], theta[:, 19]), (theta[:, 16], theta[:, 20]))
    else:
        prof = ((theta[:, 10], theta[:, 21]), (theta[:, 12], theta[:, 22]), (theta[:, 13], theta[:, 23]),
                (theta[:, 15], theta[:, 21]), (theta[:, 17], theta[:, 22]), (theta[:, 18], theta[:, 23]))
    for lat, sd in prof:
        kb = np.maximum(0, np.ceil(lat - 3 * sd))
        ke = np.maximum(kb + 1, np.floor(lat + 3 * sd))
        t.append(ke - kb + 1)
    return float(np.mean(t))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons DURING the timed regions (B200_PROFILING.md): one
    nvidia-smi process streaming a sample every 50 ms"""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                if self.stop_flag:
                    break
                self.rows.append([c.strip() for c in line.strip().split(",")])
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=2)

    def summary(self):
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        num = lambda v: float(v) if v.replace(".", "", 1).isdigit() else None
        sm = [num(r[0]) for r in rows if num(r[0]) is not None]
        mx = [num(r[1]) for r in rows if num(r[1]) is not None]
        pw = [num(r[2]) for r in rows if num(r[2]) is not None]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower().startswith("active") for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(rows)}


def cpu_baseline(ds, theta, substeps, budget_s=15.0):
    """the CPU oracle (kind 'port': R/deSolve are not installed) on a bounded sample of the
    same theta batch with every host thread"""
    from oracle import oracle
    oracle.build()
    cores = os.cpu_count() or 1
    probe = theta[: 64 * cores]
    oracle.evaluate(ds["flavour"], ds, probe[: 4 * cores], ds["period"], substeps, n_threads=cores)  # warm caches
    t0 = time.perf_counter()
    oracle.evaluate(ds["flavour"], ds, probe, ds["period"], substeps, n_threads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    n = int(min(len(theta), max(len(probe), rate * budget_s)))
    reps = max(1, int(rate * budget_s / n))  # the whole batch several times when one pass is shorter than the budget
    t0 = time.perf_counter()
    for _ in range(reps):
        oracle.evaluate(ds["flavour"], ds, theta[:n], ds["period"], substeps, n_threads=cores)
    dt = time.perf_counter() - t0
    return {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {n} proposals of the step's batch x {reps} passes, oracle/epi_oracle.c (RK4 m={substeps}), "
                      f"{cores} pthreads, {dt:.1f} s"}


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path is R + deSolve, absent
    from this image; its restatement (oracle/) is timed instead, all host threads, on a bounded
    sample of the same config.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from epi_mcmc_b200.model import load_dataset
    from oracle import oracle
    oracle.build()
    ds = load_dataset(args.workload)
    mn, mx, init = oracle.df_params(ds["flavour"], ds, ds["period"])
    cores = os.cpu_count() or 1
    probe = make_theta(args.workload, 4 * cores, 7, mn, mx, init)
    t0 = time.perf_counter()
    oracle.evaluate(ds["flavour"], ds, probe, ds["period"], args.substeps, n_threads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    total_budget = 120.0
    per_step = max(4 * cores, int(rate * total_budget / max(1, args.steps + args.warmup)))
    per_step = min(per_step, args.chains)
    theta = make_theta(args.workload, per_step, 1234, mn, mx, init)
    for _ in range(args.warmup):
        oracle.evaluate(ds["flavour"], ds, theta, ds["period"], args.substeps, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.evaluate(ds["flavour"], ds, theta, ds["period"], args.substeps, n_threads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} proposals per step (bounded sample of the {args.chains}-chain batch), {cores} pthreads"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args, ds["period"]), "chains_per_gpu": args.chains,
                       "params": len(init), "period": ds["period"], "substeps": args.substeps,
                       "theta": "init +- 0.5% of the df_params box (every chain integrates both passes)",
                       "reference_arm": "CPU restatement oracle/epi_oracle.c (R + deSolve + drjacoby are not installed)"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_name(args, period):
    what = {"A": "age-structured model-age-groups2q (BASELINE.json configs[2])",
            "I": "age-structured model-age-groups2i (secondary anchor of BASELINE.json configs[2])",
            "S": "simple SEIR model-simple-ode-drj-cmp (BASELINE.json configs[3])",
            "N": "simple SEIR + effective population size Ne (BASELINE.json configs[1])"}.get(args.workload[0], args.workload)
    return f"{args.workload}: {what}, synthetic {period}-day series, {args.chains} chains per GPU, RK4 m={args.substeps}"


_REAL_STDOUT = 1


def emit(line: dict):
    """the ONE JSON line of the contract, on the real stdout"""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


def _claim_stdout():
    """Everything else any library prints to fd 1 (NCCL's version banner, for one) goes to stderr, so that stdout
    carries exactly one JSON line.  Done when the script RUNS, not on import (tests import this module)."""
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    from epi_mcmc_b200.dist import Comm
    from epi_mcmc_b200.model import Model, load_dataset

    comm = Comm()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the evaluation path has no CPU fallback")
    dev = comm.local_rank
    torch.cuda.set_device(dev)
    ds = load_dataset(args.workload)
    M = Model(ds, substeps=args.substeps, device=dev)
    n, d = args.chains, M.d
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    theta = make_theta(args.workload, n, 1234 + comm.rank, mn, mx, init)

    # ---- device-resident arm: theta SoA already in HBM when the timed region starts
    th_dev = torch.from_numpy(np.ascontiguousarray(theta.T)).to(f"cuda:{dev}")
    logl = torch.empty(n, dtype=torch.float64, device=f"cuda:{dev}")
    logp = torch.empty_like(logl)
    status = torch.empty(n, dtype=torch.int32, device=f"cuda:{dev}")
    stream = torch.cuda.Stream(device=dev)   # the kernels are launched on THIS stream; so are the timing events
    stream.wait_stream(torch.cuda.current_stream(dev))

    def step():
        M.eval_device(th_dev.data_ptr(), n, logl.data_ptr(), logp.data_ptr(), status.data_ptr(), stream.cuda_stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize(dev)
    clocks = ClockSampler(dev)
    clocks.start()
    time.sleep(0.3)   # let nvidia-smi start streaming before the timed region
    comm.barrier()
    torch.cuda.synchronize(dev)
    launches0 = M.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    torch.cuda.synchronize(dev)
    comm.barrier()
    ms = comm.max_over_ranks(e0.elapsed_time(e1))
    launches = M.launch_count() - launches0
    value = comm.world * n * args.steps / (ms * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream, same timed loop shape)
    M.set_timing(True)
    kms = np.zeros(4)
    reps = max(3, min(args.steps, 10))
    for _ in range(reps):
        step()
        kms += np.array(M.last_kernel_ms())
    kms /= reps
    M.set_timing(False)
    torch.cuda.synchronize(dev)

    # ---- the same step on a posterior-spread cloud (secondary figure).  The headline cloud is narrow (+-0.5 % of the
    # box around init), so the chains of a warp cross every schedule breakpoint on the same day; the population of a
    # running sampler does not (its breakpoints are spread over days), which costs the ODE warps lock-step.  theta ~
    # N(mean, cov) of a 65,536-chain DE-MC population after 4,000 iterations (tests/_pop_dump.py), clipped to the box.
    cloud = None
    gpath = os.path.join(os.path.dirname(os.path.abspath(__file__)), "epi_mcmc_b200", "datasets",
                         f"{args.workload}_posterior_gauss.json")
    if os.path.exists(gpath):
        with open(gpath) as f:
            g = json.load(f)
        rng = np.random.default_rng(4321 + comm.rank)
        th_c = np.clip(rng.multivariate_normal(np.array(g["mean"]), np.array(g["cov"]), size=n), mn, mx)
        th_dev_head = th_dev
        th_dev = torch.from_numpy(np.ascontiguousarray(th_c.T)).to(f"cuda:{dev}")
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize(dev)
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(args.steps):
            step()
        c1.record(stream)
        torch.cuda.synchronize(dev)
        ms_c = comm.max_over_ranks(c0.elapsed_time(c1))
        M.set_timing(True)
        kc = np.zeros(4)
        for _ in range(reps):
            step()
            kc += np.array(M.last_kernel_ms())
        kc /= reps
        M.set_timing(False)
        torch.cuda.synchronize(dev)
        cloud = {"value": comm.world * n * args.steps / (ms_c * 1e-3), "unit": UNIT, "ms_per_step": ms_c / args.steps,
                 "kernels": {"ode_pass1_ms": float(kc[0]), "profiles_threshold_ms": float(kc[1]),
                             "ode_pass2_ms": float(kc[2]), "score_ms": float(kc[3])},
                 "status_nonzero": int((status != 0).sum().item()),
                 "theta": "N(mean, cov) of a 65,536-chain DE-MC population after 4,000 iterations, clipped to the box "
                          "(epi_mcmc_b200/datasets/%s_posterior_gauss.json): breakpoints differ by days between the "
                          "chains of a warp" % args.workload}
        th_dev = th_dev_head
        step()                      # logl / logp / status hold the headline batch again for the checks below
        torch.cuda.synchronize(dev)

    # ---- e2e arm: the public C-ABI with pinned HOST buffers, copies inside the timed region.
    # (a) epi_eval: one blocking call per step (H2D -> kernels -> D2H strictly in sequence);
    # (b) epi_eval_submit / epi_eval_wait: the same step, double-buffered — step k+1 is submitted before
    #     step k is waited for, so its input copy (copy stream) overlaps the kernels of step k.  Every
    #     step still moves its own n x d inputs from pinned host memory and its n x 20 B of results back.
    # `e2e.value` is (b), the API a host-side driver with two independent batches in flight uses;
    # (a) is reported beside it as e2e.blocking_call_value.
    from epi_mcmc_b200 import _capi
    import ctypes as C
    th_host = [torch.from_numpy(theta.copy()).pin_memory() for _ in range(2)]
    ll_host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
    lp_host = [torch.empty(n, dtype=torch.float64).pin_memory() for _ in range(2)]
    st_host = [torch.empty(n, dtype=torch.int32).pin_memory() for _ in range(2)]
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)

    def e2e_step():
        _capi.check(M._lib.epi_eval(M._h, C.cast(th_host[0].data_ptr(), dp), n, _capi.LAYOUT_AOS,
                                    C.cast(ll_host[0].data_ptr(), dp), C.cast(lp_host[0].data_ptr(), dp),
                                    C.cast(st_host[0].data_ptr(), ip)), M._h)

    def submit(k):
        b = k & 1
        _capi.check(M._lib.epi_eval_submit(M._h, b, C.cast(th_host[b].data_ptr(), dp), n, _capi.LAYOUT_AOS,
                                           C.cast(ll_host[b].data_ptr(), dp), C.cast(lp_host[b].data_ptr(), dp),
                                           C.cast(st_host[b].data_ptr(), ip)), M._h)

    def wait(k):
        _capi.check(M._lib.epi_eval_wait(M._h, k & 1), M._h)

    def pipelined(steps):
        submit(0)
        for k in range(1, steps):
            submit(k)
            wait(k - 1)
        wait(steps - 1)

    for _ in range(max(1, args.warmup)):
        e2e_step()
    comm.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()       # returns after the D2H copies completed (stream synchronised inside)
    e2e_block_s = comm.max_over_ranks(time.perf_counter() - t0)
    e2e_block_value = comm.world * n * args.steps / e2e_block_s
    assert np.allclose(ll_host[0].numpy(), logl.cpu().numpy(), rtol=0, atol=0, equal_nan=True)
    ll_host[0].zero_(); ll_host[1].zero_()
    pipelined(max(2, args.warmup))
    comm.barrier()
    t0 = time.perf_counter()
    pipelined(args.steps)
    e2e_s = comm.max_over_ranks(time.perf_counter() - t0)
    e2e_value = comm.world * n * args.steps / e2e_s
    for b in range(2):
        assert np.allclose(ll_host[b].numpy(), logl.cpu().numpy(), rtol=0, atol=0, equal_nan=True)
    clocks.stop()

    # ---- rooflines: every launch of the step against the FP64 peak measured live; `roofline` is the
    # kernel with the largest share of the step
    peak_tf = M.fp64_peak_tflops()
    fa = f_alg(ds["flavour"], ds["period"], args.substeps, mean_taps(ds["flavour"], theta), _n_terms(M),
               {"age2q": 38, "age2i": 24}.get(ds["flavour"], 3))
    # pass-2 restart: D leading days are taken from the pass-1 checkpoint instead of being integrated
    # again (bit-identical).  SURVEY §8d: report F_alg with the reduced step count too.
    off_s, pad_s, _ = M.eval_detail(theta[:4096])
    lo_ = ds["lockdown_offset"]
    if ds["flavour"] == "age2q":
        minbp = off_s + lo_ + theta[:4096, 18]
        cap = 59
    elif ds["flavour"] == "age2i":
        minbp = off_s + lo_ + theta[:4096, 20]
        cap = min(ds["period"], 64) - 1
    else:
        minbp = (off_s + lo_).astype(float)
        cap = min(ds["period"], 64) - 1
    restart_on = os.environ.get("EPI_PASS2_RESTART", "1") != "0"
    D = np.clip(np.floor(minbp) - (pad_s + 1), 0, cap) if restart_on else np.zeros(len(off_s))
    D = np.where(off_s == 10000, 0, D)
    d_mean = float(D.mean())
    # data windows: calclogl reads the model series only on [dstart, dstart + n - 1] of every scored series
    # (window clamps of model-age-groups2q.R:497-512), so the scoring path integrates pass 2 up to the last
    # window day and convolves window days (+ one predecessor) only — "convolving only observed days" of
    # SURVEY §8d's list of admissible reductions.  Same window arithmetic as the kernels, in numpy.
    P_ = ds["period"]
    n_obs = ([ds["n_dcasei"], ds["n_dcasei"], ds["n_dhospi"], ds["n_dmorti"], ds["n_dmorti"]]
             if ds["flavour"] in ("age2q", "age2i") else [ds["n_dhospi"], ds["n_dmorti"]])
    off_e = np.where(off_s == 10000, 1, off_s).astype(np.int64)
    T_ = pad_s.astype(np.int64) + P_
    tlo, thi = T_.copy(), np.ones_like(T_)
    for nn in n_obs:
        de = off_e + nn - 1
        dsr = np.where(de > T_, T_ - nn + 1, off_e)
        dsr = np.maximum(dsr, 1)
        tlo, thi = np.minimum(tlo, dsr), np.maximum(thi, dsr + nn - 1)
    nint = np.clip(thi - pad_s, 1, P_)                                   # pass-2 days integrated (incl. day 0)
    swept = np.minimum(thi - (pad_s + 1), P_ - 1) - (tlo - 1 - (pad_s + 1)) + 1  # days the score sweep covers
    nint = np.where(off_s == 10000, D, nint)                             # invalid offset: pass 2 is skipped altogether
    days_pass2 = float(np.maximum(nint - D, 0).mean())
    days_swept = float(np.minimum(swept, P_ + 64).mean())
    step_flop = fa["per_kernel"]["ode_pass2"] / P_                        # per integrated day
    # whole-period pass 1 (simple / Ne / age-2i) in two rounds: 128 days for everybody, all days again for the
    # chains that did not cross the threshold there (kPass1Chunk in epi_kernels.cu)
    P1_ = 60 if ds["flavour"] == "age2q" else P_
    chunk_on = os.environ.get("EPI_PASS1_CHUNK", "1") != "0" and P1_ > 160
    if chunk_on:
        first = off_s + lo_ - pad_s - 1
        retry = (off_s == 10000) | (first >= 128)
        days_pass1 = float(128 + retry.mean() * P1_)
    else:
        days_pass1 = float(P1_)
    r_share = fa["ode_executed_share"] if os.environ.get("EPI_FORCE_R_FALLBACK", "0") != "1" else 1.0
    fa["per_kernel_executed"] = dict(fa["per_kernel"], ode_pass1=fa["per_kernel"]["ode_pass1"] * days_pass1 / P1_ * r_share,
                                     ode_pass2=days_pass2 * step_flop * r_share,
                                     score=fa["per_kernel"]["score"] - (P_ - min(days_swept, P_)) * fa["score_per_day"])
    fa["total_executed"] = sum(fa["per_kernel_executed"].values())
    knames = ["ode_pass1", "profiles_threshold", "ode_pass2", "score"]
    kernel_fn = {"ode_pass1": "k_ode<pass1>", "profiles_threshold": "k_mid", "ode_pass2": "k_ode<pass2>", "score": "k_score"}
    per_kernel = {}
    ncu_match = (args.workload, n, args.substeps) == NCU_WORKLOAD
    for k, msk in zip(knames, kms):
        tf = fa["per_kernel_executed"][k] * n / (float(msk) * 1e-3) / 1e12
        per_kernel[k] = {"kernel": kernel_fn[k], "ms": float(msk),
                         "algorithmic_flop_per_launch": fa["per_kernel_executed"][k] * n,
                         "algorithmic_flop_per_launch_without_reductions": fa["per_kernel"][k] * n,
                         "achieved": tf, "frac": tf / peak_tf if peak_tf else None,
                         "traffic": NCU_TRAFFIC.get(k) if ncu_match else None,
                         "fp64_pipe_pct_ncu": NCU_FP64_PIPE_PCT.get(k) if ncu_match else None,
                         "executed_fp64_flop_per_launch_ncu": NCU_EXECUTED_FLOP.get(k) if ncu_match else None,
                         "executed_frac": (NCU_EXECUTED_FLOP[k] / (float(msk) * 1e-3) / 1e12 / peak_tf
                                           if ncu_match and peak_tf else None)}
    dom = knames[int(np.argmax(kms))]
    step_ms = ms / args.steps
    # the HBM view of the same launches (why the bound is FP64 and not memory): ALGORITHMIC bytes per launch — what
    # each kernel has to read and write once — against the measured copy bandwidth of MEASURED_PEAKS.json
    hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        pass
    age = ds["flavour"] in ("age2q", "age2i")
    NGb, NEQb, padb = (2, 8, 80 if ds["flavour"] == "age2i" else 64) if age else (1, 3, 64)
    ck_days = min(P1_, 64)
    row = lambda days: NGb * ((days + 1) & ~1) * 8
    bytes_alg = {
        "ode_pass1": d * 8 + row(P1_) + NEQb * ck_days * 8,                       # theta in; S rows + checkpoints out
        "profiles_threshold": d * 8 + row(P1_) + NGb * padb * 4 * 8 + 6 * 8 + 12,  # theta, pass-1 S in; profile table, meta out
        "ode_pass2": d * 8 + NEQb * 8 + row(min(int(d_mean) + 1, P1_)) + row(P_),  # theta, one checkpoint, head of pass-1 S in; S rows out
        "score": d * 8 + row(P_) + NGb * padb * 4 * 8 + 20,                       # theta, S rows, profile table in; logl, logp, status out
    }
    for k, msk in zip(knames, kms):
        gbs = bytes_alg[k] * n / (float(msk) * 1e-3) / 1e9
        per_kernel[k]["hbm"] = {"algorithmic_bytes_per_launch": bytes_alg[k] * n, "achieved": gbs, "peak": hbm_peak,
                                "unit": "GB/s", "frac": gbs / hbm_peak}
    whole_tf = fa["total_executed"] * n / (step_ms * 1e-3) / 1e12

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": comm.world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(args, ds["period"]),
                   "chains_per_gpu": n, "params": d, "period": ds["period"], "substeps": args.substeps,
                   "theta": "init +- 0.5% of the df_params box (every chain integrates both passes)",
                   "l2": "inputs larger than L2: the S-history/profile working set of one step is "
                         f"{_ws_mb(ds, n):.0f} MB vs 126 MB L2",
                   "parallelism": f"chains sharded over {comm.world} GPU(s), no data-path collective"},
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * d * 8), "d2h_bytes_per_step": int(n * 20),
                "api": "epi_eval_submit/epi_eval_wait, two batches in flight (input copy of step k+1 overlaps the kernels of step k)",
                "blocking_call_value": e2e_block_value, "blocking_call_api": "epi_eval, one blocking call per step"},
        "gpu_launches": int(launches),
        "kernels": {"ode_pass1_ms": float(kms[0]), "profiles_threshold_ms": float(kms[1]), "ode_pass2_ms": float(kms[2]),
                    "score_ms": float(kms[3])},
        "roofline": {"bound": "fp64", "kernel": per_kernel[dom]["kernel"], "achieved": per_kernel[dom]["achieved"],
                     "peak": peak_tf, "unit": "TFLOP/s", "frac": per_kernel[dom]["frac"],
                     "traffic": per_kernel[dom]["traffic"],
                     "peak_source": "measured live: DFMA-chain microbenchmark epi_fp64_peak (MEASURED_PEAKS.json has no FP64 entry; "
                                    "nominal 37.2 TFLOP/s = 148 SM x 64 DFMA/clk x 2 x 1.965 GHz)",
                     "algorithmic_flop_per_launch": per_kernel[dom]["algorithmic_flop_per_launch"],
                     "hbm": dict(per_kernel[dom]["hbm"], peak_source=hbm_src),
                     "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, "
                                       "profiles/r1_v15_ncu_full.txt (65,536 chains, A300)"},
        "roofline_all": per_kernel,
        "posterior_cloud": cloud,
        "f_alg": {"per_eval": fa["total"], "per_eval_executed": fa["total_executed"],
                  "pass2_restart_days_mean": d_mean, "pass2_days_integrated_mean": days_pass2,
                  "score_days_swept_mean": days_swept, "pass1_days_integrated_mean": days_pass1, "period": P_, "ode_share": fa["ode"] / fa["total"],
                  "whole_step_tflops": whole_tf, "whole_step_frac": whole_tf / peak_tf if peak_tf else None,
                  "whole_step_frac_if_counted_without_reductions":
                      (fa["total"] * n / (step_ms * 1e-3) / 1e12) / peak_tf if peak_tf else None,
                  "whole_step_executed_frac_ncu": (sum(NCU_EXECUTED_FLOP.values()) / (step_ms * 1e-3) / 1e12 / peak_tf
                                                   if ncu_match and peak_tf else None),
                  "note": "per_eval is the SURVEY 8d figure (the reference's arithmetic; the ODE kernels execute about a "
                          "third of it, so their algorithmic frac exceeds 1 — executed_frac / fp64_pipe_pct_ncu in "
                          "roofline_all are counted from the ncu SASS table instead); per_eval_executed and every frac count only the work "
                          "actually done: pass 2 resumes from the pass-1 checkpoint after pass2_restart_days_mean "
                          "bit-identical days and stops at the last data-window day, the score sweep covers the data "
                          "windows plus one predecessor day, and the removed compartment R is not integrated (it only "
                          "enters the reference through a bounds guard that provably cannot fire on it; chains where that "
                          "is not certain are recomputed with R) — results identical bit for bit; DESIGN.md section 4"},
    }
    if comm.rank == 0 and comm.world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(ds, theta, args.substeps)
    if not args.no_sampler:
        line["sampler"] = sampler_leg(M, comm, args)
    if comm.rank == 0:
        emit(line)
    M.close()
    comm.close()


def _n_terms(M):
    # rows of the likelihood term table (a-7): age-2q with n obs days: 2(n+2) + (n+10) + 1 + 2n
    ds = M.data
    if ds["flavour"] == "age2q":
        n = ds["n_dcasei"]
        return 2 * (n + 2) + (ds["n_dhospi"] + 10) + 1 + 2 * ds["n_dmorti"]
    if ds["flavour"] == "age2i":
        return 2 * (ds["n_dcasei"] + 2) + ds["n_dhospi"] + 2 * ds["n_dmorti"]
    return ds["n_dhospi"] + ds["n_dmorti"] + 1


def _ws_mb(ds, n):
    P = ds["period"]
    if ds["flavour"] == "age2q":
        return n * ((P + 60) * 2 * 8 + 6 * 64 * 8) / 1e6
    if ds["flavour"] == "age2i":
        return n * (2 * P * 2 * 8 + 2 * 80 * 4 * 8) / 1e6
    return n * (2 * P * 8 + 2 * 64 * 8) / 1e6


def sampler_leg(M, comm, args):
    """DE-MC with the fused propose/accept kernel: chain state stays on the GPU; the population
    is all-gathered over NCCL every 10 iterations.  evals/s inside the sampler, and ESS/s over the
    recorded sampling phase (coda spectral ESS of `fitkeyparamnames`, measured on 64 chains of this
    rank and scaled to the whole population; min and median over the parameters)."""
    import torch
    from epi_mcmc_b200.sampler import Sampler, pooled_ess
    n = args.chains
    S = Sampler(M, n, kind="demc", seed=42, chain_id0=comm.rank * n, subspace=not args.sampler_fulldim)
    comm.attach(S)
    burn, iters = args.sampler_burnin, args.sampler_iters
    S.adapt(True)
    torch.cuda.synchronize()
    comm.barrier()
    tb = time.perf_counter()
    for _ in range(burn // 10):
        S.run(10)
        comm.exchange(S)
    S.adapt(False)
    S.record(1, iters)
    torch.cuda.synchronize()
    comm.barrier()
    burn_s = comm.max_over_ranks(time.perf_counter() - tb)
    ev0 = S.stats()["evals"]
    t0 = time.perf_counter()
    for _ in range(iters // 10):
        S.run(10)
        comm.exchange(S)
    torch.cuda.synchronize()
    dt = comm.max_over_ranks(time.perf_counter() - t0)
    st = S.stats()
    n_sub = min(64, n)
    draws, _ = S.draws_subset(0, n_sub)
    keys = [M.fit_paramnames.index(k) for k in M.fitkeyparamnames if k in M.fit_paramnames]
    ess = pooled_ess(draws[:, :, keys]) * (n / n_sub)
    return {"kind": "DE-MC (scale mixture, subspace updates), population snapshot all-gathered every 10 iterations",
            "chains_per_gpu": n, "burnin_iterations": burn, "burnin_s": burn_s, "sampling_iterations": iters,
            "sampling_s": dt, "evals_per_s": comm.world * (st["evals"] - ev0) / dt, "accept_rate": st["accept_rate"],
            "ess_per_s_min": comm.world * float(ess.min()) / dt, "ess_per_s_median": comm.world * float(np.median(ess)) / dt,
            "ess_per_eval_median": float(np.median(ess)) / max(1, st["evals"] - ev0),
            "note": "ESS of 64 chains of rank 0 scaled to all chains and ranks; sampling phase only; chains start "
                    "at the model file's init + 2% box jitter.  The spectral ESS of a chain shorter than a few "
                    "integrated autocorrelation times is biased upwards: on this target 200 + 400 iterations report "
                    "~158 k ESS/s per GPU, 1000 + 3000 ~25 k, 2000 + 6000 ~12 k (profiles/r1_summary.md)"}


if __name__ == "__main__":
    main()
