#!/usr/bin/env python
"""bench.py — log-posterior evals/s of the epi-mcmc hot path on N B200s.

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference --gpus N --steps K --warmup W

One "step" = one pass of the hot path (calclogl(transformParams(theta)) + calclogp(theta)) over one batch of
proposals.  Workload = BASELINE.json configs[2]: the age-structured model (model-age-groups2q) on the synthetic A300
series, RK4 m = 4, `--chains` = 65,536 chains IN TOTAL: under `--scaling strong` (default, the configuration
BASELINE.json quotes: "65,536 chains at 1/2/4/8 B200") every GPU evaluates 65,536 / N of them; `--scaling weak` gives
every GPU all 65,536 (also reported as the `weak` object of a strong run).  Chains are independent: no data-path
collective.  The headline theta are draws from the posterior cloud of a running sampler (what the kernels see in
production: breakpoints differ by days between the chains of a warp); the narrow cloud of round 1 is `narrow_cloud`.

JSON line keys beyond the base contract:
  roofline      the longest kernel of the step: EXECUTED FP64 flop (ncu SASS counts of this workload,
                profiles/ncu_constants.json) / launch time / the DFMA peak measured live — never above 1; the SURVEY 8d
                formula figure is `survey_formula_frac`; `hbm` is the HBM view of the same launch
  cpu_baseline  the CPU oracle ("port": R is probed for and not installed) on a bounded sample, all host threads,
                and a CPU Metropolis sampler on it for ESS/s
  e2e           the same metric through the host-buffer C ABI (epi_eval_submit/wait, and epi_eval beside it)
                with pinned HOST buffers (H2D + D2H inside)
  sustained     the same step for >= 1 s
  kernels       mean device ms of the four launches of one step
  secondary     one short timed leg per other BASELINE config (S120 / S365 at 2^20 proposals, A365, NE120 at 4,096,
                A300 on the daily grid m = 1, I290)
  sampler       fused propose/accept DE-MC run: evals/s inside the sampler, ESS/s from independent replicate
                populations, split R-hat, and `k2`: the fused kernel alone against the HBM peak (65,536 and 2^22 chains)
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

# ncu constants of bench.py's own workloads (profiles/extract_ncu_constants.py writes the file from a
# `ncu --set full` capture): DRAM bytes per launch, FP64 pipe utilisation, FP64 flop actually EXECUTED per launch
# (thread-level predicated-on DFMA x 2 + DMUL + DADD of the report's SASS table).  SURVEY 8d counts the reference's
# arithmetic (bounds-guard comparisons, the output variables Re/Rt, the removed compartment, per-call segment
# selection ...): the ODE kernels execute about a third of that, so the executed count is the honest numerator.
def load_ncu_constants():
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_constants.json")) as f:
            return json.load(f)
    except (OSError, ValueError):
        return {}


METRIC = "log_posterior_evals_per_s"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)   # ~1 s of timed region at 3 ms per step
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="A300")
    ap.add_argument("--chains", type=int, default=65536, help="chains of the configuration IN TOTAL (BASELINE.json: 65,536)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: every GPU takes chains / N (BASELINE.json configs[2]); weak: every GPU takes all of them")
    ap.add_argument("--substeps", type=int, default=4)
    ap.add_argument("--theta", default="auto", choices=["auto", "posterior", "narrow"],
                    help="headline proposals: posterior cloud of a running sampler (auto: when the workload has one) or init +- 0.5 %% of the box")
    ap.add_argument("--no-sampler", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--sampler-iters", type=int, default=3000)
    ap.add_argument("--sampler-burnin", type=int, default=4000)
    ap.add_argument("--sampler-fulldim", action="store_true", help="DE-MC without subspace (crossover) updates")
    return ap.parse_args()


def make_theta(ds_name, n, seed, mn, mx, init):
    """narrow cloud: +-0.5 % of the box width around the model file's init (= the mode of the synthetic data), so
    every chain does the full work (valid data_offset, pass 2 integrated).  Box-uniform draws would let ~1/3 of the
    chains skip pass 2 and flatter the number."""
    rng = np.random.default_rng(seed)
    th = init + (rng.uniform(size=(n, len(init))) - 0.5) * 0.01 * (mx - mn)
    return np.clip(th, mn, mx)


def posterior_gauss(workload):
    """mean and covariance of a 65,536-chain DE-MC population after 4,000 iterations (tests/_pop_dump.py), or None"""
    path = os.path.join(ROOT, "epi_mcmc_b200", "datasets", f"{workload}_posterior_gauss.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        g = json.load(f)
    return np.array(g["mean"]), np.array(g["cov"])


def make_cloud(workload, n, seed, mn, mx):
    """posterior cloud: theta ~ N(mean, cov) of a running sampler's population, clipped to the box.  The breakpoints
    of the chains of a warp then differ by days, which costs the ODE warps their lock-step."""
    g = posterior_gauss(workload)
    rng = np.random.default_rng(seed)
    return np.clip(rng.multivariate_normal(g[0], g[1], size=n), mn, mx)


THETA_LABEL = {"posterior": "N(mean, cov) of a 65,536-chain DE-MC population after 4,000 iterations, clipped to the box "
                            "(epi_mcmc_b200/datasets/<workload>_posterior_gauss.json): what a running sampler evaluates",
               "narrow": "init +- 0.5% of the df_params box (every chain integrates both passes, warps in lock-step)"}


def batch_theta(kind, workload, total, mn, mx, init):
    """the configuration's whole batch of `total` proposals — the same on every rank, whatever the sharding"""
    return make_cloud(workload, total, 4321, mn, mx) if kind == "posterior" else make_theta(workload, total, 1234, mn, mx, init)


def r_probe():
    """BASELINE.md step 1 / SURVEY 8d: is the reference's own runtime here?  The reference arm can only be R when
    Rscript, deSolve AND the reference tree are all present (the tree never travels to the GPU box)."""
    import shutil
    rs = shutil.which("Rscript")
    have_ds = None
    if rs:
        try:
            have_ds = subprocess.run([rs, "-e", "quit(status = !requireNamespace('deSolve', quietly = TRUE))"],
                                     capture_output=True, timeout=60).returncode == 0
        except Exception:
            have_ds = False
    return {"Rscript": rs, "deSolve": have_ds, "reference_tree": os.path.isdir("/root/reference"),
            "usable": bool(rs and have_ds and os.path.isdir("/root/reference"))}


def f_alg(flavour, P, m, taps_mean, n_terms, n_prior):
    """ALGORITHMIC FLOP per evaluation, SURVEY.md §8d"""
    age = flavour in ("age2q", "age2i", "age2x")
    S, K, C1, C2 = (10, 6, 2, 6) if age else (4, 2, 1, 2)
    P1 = 60 if flavour in ("age2q", "age2x") else P  # pass 1: 60 days in 2q / 2x, the whole period in 2i / simple
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
    n_r = 0 if flavour == "age2x" else 2 if age else 1   # (2x: waning immunity, R is always integrated)
    return dict(total=total, ode=ode, per_kernel=per_kernel, score_per_day=C2 * 2 * taps_mean + 6 * C2,
                ode_executed_share=(4 * F_rhs + 14 * (S - n_r)) / (4 * F_rhs + 14 * S))


def mean_taps(flavour, theta):
    if flavour not in ("age2q", "age2i", "age2x"):
        return 16.0
    t = []
    if flavour == "age2x":
        prof = ((theta[:, 46], 2.5), (theta[:, 11], theta[:, 19]), (theta[:, 12], theta[:, 20]),
                (theta[:, 47], 2.5), (theta[:, 15], theta[:, 19]), (theta[:, 16], theta[:, 20]))
    elif flavour == "age2q":
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


def cpu_baseline(ds, theta, substeps, budget_s=15.0, ess_budget_s=20.0, workload=None):
    """the CPU oracle (kind 'port': R/deSolve are probed for and not installed) on a bounded sample of the
    same theta batch with every host thread; then a CPU Metropolis sampler on the same density (oracle/cpu_mcmc.py,
    one chain per host thread, proposal shape = the posterior covariance) for ESS/s"""
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
    out = {"value": n * reps / dt, "unit": UNIT, "cores": cores, "kind": "port", "r_probe": r_probe(),
           "sample": f"first {n} proposals of the step's batch x {reps} passes, oracle/epi_oracle.c (RK4 m={substeps}), "
                     f"{cores} pthreads, {dt:.1f} s"}
    g = posterior_gauss(workload) if workload else None
    if g is not None and ess_budget_s > 0:
        from oracle import cpu_mcmc
        from epi_mcmc_b200.sampler import effective_size
        mn, mx, _ = (np.asarray(v) for v in oracle.df_params(ds["flavour"], ds, ds["period"]))
        nc = max(4, cores)
        x0 = np.clip(np.random.default_rng(11).multivariate_normal(g[0], g[1], size=nc), mn, mx)
        draws, info = cpu_mcmc.run(oracle, ds, substeps, nc, 100, 10 ** 9, 1, 5, n_threads=cores, cov0=g[1], x0=x0,
                                   budget_s=ess_budget_s)
        # independent chains started in the stationary law: coda's spectral ESS per chain, summed over the chains
        ess = np.array([sum(effective_size(draws[:, c, p]) for c in range(nc)) for p in range(draws.shape[2])])
        out["ess_per_s_median"] = float(np.median(ess)) / info["sampling_s"]
        out["ess_per_s_min"] = float(ess.min()) / info["sampling_s"]
        out["ess_sample"] = (f"{nc} independent random-walk Metropolis chains on the oracle density, started from draws of the "
                             f"posterior's Gaussian summary with its covariance x 2.38^2/d as the proposal shape, "
                             f"{info['sampling_iterations']} iterations each in {info['sampling_s']:.1f} s "
                             f"({info['evals_sampling'] / info['sampling_s']:.0f} evals/s, acceptance {info['accept_rate']:.3f}); "
                             f"ESS = coda spectral ESS per chain, summed; median / min over all {draws.shape[2]} parameters")
    return out


def make_config(args, period, n_per_gpu, world, d, theta_kind, ws_mb=None):
    """the `config` object — the SAME keys and values in both arms, so that the driver can compare them"""
    return {"workload": workload_name(args, period), "chains_total": args.chains, "chains_per_gpu": n_per_gpu,
            "params": d, "period": period, "substeps": args.substeps, "theta": THETA_LABEL[theta_kind],
            "l2": "inputs larger than L2: the S-history/profile working set of one step is "
                  f"{ws_mb:.0f} MB vs 126 MB L2" if ws_mb is not None else "n/a",
            "parallelism": f"{args.scaling} scaling: {args.chains} chains "
                           + ("sharded over" if args.scaling == "strong" else "on each of") + f" {world} GPU(s), no data-path collective"}


def theta_kind_of(args):
    if args.theta == "auto":
        return "posterior" if posterior_gauss(args.workload) is not None else "narrow"
    return args.theta


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path is R + deSolve (probed for, absent from
    this image); its restatement (oracle/) is timed instead, all host threads, on a bounded sample of the same
    config.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from epi_mcmc_b200.model import load_dataset
    from oracle import oracle
    oracle.build()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    ds = load_dataset(args.workload)
    mn, mx, init = oracle.df_params(ds["flavour"], ds, ds["period"])
    cores = os.cpu_count() or 1
    kind = theta_kind_of(args)
    n_per_gpu = (args.chains + world - 1) // world if args.scaling == "strong" else args.chains
    batch = batch_theta(kind, args.workload, min(args.chains, 16384), mn, mx, init)
    probe = batch[: 4 * cores]
    t0 = time.perf_counter()
    oracle.evaluate(ds["flavour"], ds, probe, ds["period"], args.substeps, n_threads=cores)
    rate = len(probe) / (time.perf_counter() - t0)
    total_budget = 120.0
    per_step = max(4 * cores, int(rate * total_budget / max(1, args.steps + args.warmup)))
    per_step = min(per_step, len(batch))
    theta = batch[:per_step]
    for _ in range(args.warmup):
        oracle.evaluate(ds["flavour"], ds, theta, ds["period"], args.substeps, n_threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.evaluate(ds["flavour"], ds, theta, ds["period"], args.substeps, n_threads=cores)
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    sample = f"{per_step} proposals per step (bounded sample of the {args.chains}-chain batch), {cores} pthreads"
    ws = _ws_mb(ds, n_per_gpu)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(args, ds["period"], n_per_gpu, world, len(init), kind, ws),
            "reference_arm": "CPU restatement oracle/epi_oracle.c (R + deSolve + drjacoby are not installed: r_probe)",
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "r_probe": r_probe()},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_name(args, period):
    what = {"A": "age-structured model-age-groups2q (BASELINE.json configs[2])",
            "I": "age-structured model-age-groups2i (secondary anchor of BASELINE.json configs[2])",
            "S": "simple SEIR model-simple-ode-drj-cmp (BASELINE.json configs[3])",
            "N": "simple SEIR + effective population size Ne (BASELINE.json configs[1])"}.get(args.workload[0], args.workload)
    return f"{args.workload}: {what}, synthetic {period}-day series, {args.chains} chains in total, RK4 m={args.substeps}"


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


class DeviceBatch:
    """one device-resident batch of this rank: theta SoA in HBM, result buffers, the launching stream"""

    def __init__(self, M, theta, dev, stream):
        import torch
        self.M, self.n, self.stream = M, len(theta), stream
        self.th = torch.from_numpy(np.ascontiguousarray(theta.T)).to(f"cuda:{dev}")
        self.logl = torch.empty(self.n, dtype=torch.float64, device=f"cuda:{dev}")
        self.logp = torch.empty_like(self.logl)
        self.status = torch.empty(self.n, dtype=torch.int32, device=f"cuda:{dev}")

    def step(self):
        self.M.eval_device(self.th.data_ptr(), self.n, self.logl.data_ptr(), self.logp.data_ptr(), self.status.data_ptr(),
                           self.stream.cuda_stream)


def timed(batch, steps, warmup, comm, dev):
    """W untimed steps, then EXACTLY `steps` steps between CUDA events recorded on the launching stream, bracketed
    by a barrier + synchronize on both sides; max over ranks.  Returns (ms total, launches)."""
    import torch
    for _ in range(warmup):
        batch.step()
    torch.cuda.synchronize(dev)
    comm.barrier()
    torch.cuda.synchronize(dev)
    l0 = batch.M.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(batch.stream)
    for _ in range(steps):
        batch.step()
    e1.record(batch.stream)
    torch.cuda.synchronize(dev)
    comm.barrier()
    return comm.max_over_ranks(e0.elapsed_time(e1)), batch.M.launch_count() - l0


def kernel_ms(batch, reps, dev):
    """median device time of the four launches of a step over `reps` steps (CUDA events on the launching stream)"""
    import torch
    M = batch.M
    M.set_timing(True)
    k = []
    for _ in range(reps):
        batch.step()
        k.append(M.last_kernel_ms())
    M.set_timing(False)
    torch.cuda.synchronize(dev)
    return np.median(np.array(k), axis=0)   # (median: one disturbed repetition must not colour the per-kernel split)


KNAMES = ["ode_pass1", "profiles_threshold", "ode_pass2", "score"]
KERNEL_FN = {"ode_pass1": "k_ode<pass1>", "profiles_threshold": "k_mid", "ode_pass2": "k_ode<pass2>", "score": "k_score"}


def kdict(k):
    return {"ode_pass1_ms": float(k[0]), "profiles_threshold_ms": float(k[1]), "ode_pass2_ms": float(k[2]), "score_ms": float(k[3])}


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def rooflines(M, ds, args_substeps, theta, n, kms, step_ms, peak_tf, ncu):
    """FP64 and HBM view of every launch of the step.  `frac` is EXECUTED flop (ncu SASS count per chain of this
    workload, when profiles/ncu_constants.json has it) / time / DFMA peak; `survey_formula_frac` prices the launch
    with SURVEY 8d's formula (the reference's arithmetic, reductions applied) and can exceed 1 for the ODE launches."""
    d = M.d
    fa = f_alg(ds["flavour"], ds["period"], args_substeps, mean_taps(ds["flavour"], theta), _n_terms(M),
               {"age2q": 38, "age2i": 24, "age2x": 41}.get(ds["flavour"], 3))
    # pass-2 restart: D leading days are taken from the pass-1 checkpoint instead of being integrated
    # again (bit-identical).  SURVEY §8d: report F_alg with the reduced step count too.
    off_s, pad_s, _ = M.eval_detail(theta[:4096])
    th = theta[:4096]
    lo_ = ds["lockdown_offset"]
    if ds["flavour"] in ("age2q", "age2x"):
        minbp = off_s + lo_ + th[:, 18]
        cap = 59
    elif ds["flavour"] == "age2i":
        minbp = off_s + lo_ + th[:, 20]
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
             if ds["flavour"] in ("age2q", "age2i", "age2x") else [ds["n_dhospi"], ds["n_dmorti"]])
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
    P1_ = 60 if ds["flavour"] in ("age2q", "age2x") else P_
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
    hbm, hbm_src = hbm_peak()
    age = ds["flavour"] in ("age2q", "age2i", "age2x")
    NGb, NEQb, padb = (2, 10 if ds["flavour"] == "age2x" else 8, 80 if ds["flavour"] == "age2i" else 64) if age else (1, 3, 64)
    ck_days = min(P1_, 64)
    row = lambda days: NGb * ((days + 1) & ~1) * 8
    bytes_alg = {
        "ode_pass1": d * 8 + row(P1_) + NEQb * ck_days * 8,                       # theta in; S rows + checkpoints out
        "profiles_threshold": d * 8 + row(P1_) + NGb * padb * 4 * 8 + 6 * 8 + 12,  # theta, pass-1 S in; profile table, meta out
        "ode_pass2": d * 8 + NEQb * 8 + row(min(int(d_mean) + 1, P1_)) + row(P_),  # theta, one checkpoint, head of pass-1 S in; S rows out
        "score": d * 8 + row(P_) + NGb * padb * 4 * 8 + 20,                       # theta, S rows, profile table in; logl, logp, status out
    }
    per_kernel = {}
    kc = (ncu or {}).get("kernels", {})
    n_cap = (ncu or {}).get("chains", 0)
    for k, msk in zip(KNAMES, kms):
        t = float(msk) * 1e-3
        tf_formula = fa["per_kernel_executed"][k] * n / t / 1e12
        c = kc.get(k)
        # executed flop per chain from the capture of this workload (the count per chain does not depend on the batch)
        ex = c["executed_fp64_flop"] / n_cap * n if c and n_cap else None
        tf = ex / t / 1e12 if ex else None
        gbs = bytes_alg[k] * n / t / 1e9
        per_kernel[k] = {"kernel": KERNEL_FN[k], "ms": float(msk),
                         "executed_fp64_flop_per_launch": ex, "achieved": tf, "frac": (tf / peak_tf if tf and peak_tf else None),
                         "fp64_pipe_pct_ncu": c["fp64_pipe_pct"] if c and n == n_cap else None,
                         "traffic": c["traffic_bytes"] if c and n == n_cap else None,
                         "survey_formula_flop_per_launch": fa["per_kernel_executed"][k] * n,
                         "survey_formula_flop_per_launch_without_reductions": fa["per_kernel"][k] * n,
                         "survey_formula_tflops": tf_formula,
                         "survey_formula_frac": tf_formula / peak_tf if peak_tf else None,
                         "hbm": {"algorithmic_bytes_per_launch": bytes_alg[k] * n, "achieved": gbs, "peak": hbm, "unit": "GB/s",
                                 "frac": gbs / hbm, "peak_source": hbm_src}}
    ex_tot = sum(v["executed_fp64_flop_per_launch"] for v in per_kernel.values()) if all(
        v["executed_fp64_flop_per_launch"] for v in per_kernel.values()) else None
    whole_tf = fa["total_executed"] * n / (step_ms * 1e-3) / 1e12
    fsum = {"per_eval_survey_formula": fa["total"], "per_eval_survey_formula_with_reductions": fa["total_executed"],
            "per_eval_executed_ncu": ex_tot / n if ex_tot else None,
            "pass2_restart_days_mean": d_mean, "pass2_days_integrated_mean": days_pass2,
            "score_days_swept_mean": days_swept, "pass1_days_integrated_mean": days_pass1, "period": P_,
            "whole_step_executed_tflops": ex_tot / (step_ms * 1e-3) / 1e12 if ex_tot else None,
            "whole_step_executed_frac": ex_tot / (step_ms * 1e-3) / 1e12 / peak_tf if ex_tot and peak_tf else None,
            "whole_step_survey_formula_frac": whole_tf / peak_tf if peak_tf else None,
            "whole_step_survey_formula_frac_without_reductions":
                (fa["total"] * n / (step_ms * 1e-3) / 1e12) / peak_tf if peak_tf else None,
            "note": "executed = FP64 flop the kernels actually execute (ncu SASS counts, profiles/ncu_constants.json); the "
                    "survey formula prices the reference's arithmetic (124 flop per RHS call incl. guard comparisons, output "
                    "variables, the removed compartment), about three times what the ODE kernels execute, so its fractions "
                    "are NOT roofline fractions.  Reductions (all bit-identical, DESIGN.md section 4): pass 2 resumes from the "
                    "pass-1 checkpoint and stops at the last data-window day, the score sweep covers the data windows, the "
                    "removed compartment R is not integrated"}
    return per_kernel, fsum


def device_leg(M, ds, theta, args, comm, dev, stream, steps, warmup, peak_tf, ncu, world_chains):
    """device-resident arm on one theta batch: value, per-kernel times, rooflines"""
    b = DeviceBatch(M, theta, dev, stream)
    ms, launches = timed(b, steps, warmup, comm, dev)
    kms = kernel_ms(b, max(3, min(steps, 10)), dev)
    step_ms = ms / steps
    pk, fsum = rooflines(M, ds, args.substeps, theta, len(theta), kms, step_ms, peak_tf, ncu)
    return {"value": world_chains * steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": step_ms, "steps": steps,
            "kernels": kdict(kms), "status_nonzero": int((b.status != 0).sum().item())}, b, kms, launches, pk, fsum


def secondary_legs(args, comm, dev, stream, peak_tf, ncu_all):
    """one short timed leg per other BASELINE config, sharded like the headline (strong: total / N per GPU)"""
    from epi_mcmc_b200.model import Model, load_dataset
    import torch
    legs = [("S120", 1 << 20, 4, "configs[0]/[3]: simple SEIR, 120-day horizon, 2^20 proposals"),
            ("S365", 1 << 20, 4, "configs[3]: simple SEIR, 365-day horizon, 2^20 proposals"),
            ("NE120", 4096, 4, "configs[1]: model-Ne, 4,096 chains"),
            ("A365", 65536, 4, "configs[2] at the 365-day horizon"),
            ("A300", 65536, 1, "configs[2] on the daily grid (m = 1, the north_star's wording)"),
            ("I290", 65536, 4, "configs[2], secondary anchor model-age-groups2i"),
            ("X300", 65536, 4, "configs[2], later generation model-age-groups2x (waning immunity, incidence-based observation model)")]
    out = []
    for wl, total, m, what in legs:
        try:
            ds = load_dataset(wl)
        except OSError:
            continue
        n = (total + comm.world - 1) // comm.world if args.scaling == "strong" else total
        M = Model(ds, substeps=m, device=dev)
        mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
        kind = "posterior" if (posterior_gauss(wl) is not None and args.theta != "narrow") else "narrow"
        whole = batch_theta(kind, wl, total, mn, mx, init)
        theta = whole[comm.rank * n:(comm.rank + 1) * n] if args.scaling == "strong" else whole
        b = DeviceBatch(M, theta, dev, stream)
        for _ in range(3):
            b.step()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter(); b.step(); torch.cuda.synchronize(dev)
        est = max(1e-4, time.perf_counter() - t0)
        steps = int(min(200, max(5, 0.25 / est)))
        ms, _ = timed(b, steps, 3, comm, dev)
        kms = kernel_ms(b, 3, dev)
        ncu = ncu_all.get(f"{wl}/m{m}/{kind}")
        pk, fsum = rooflines(M, ds, m, theta, len(theta), kms, ms / steps, peak_tf, ncu)
        tot_chains = total if args.scaling == "strong" else total * comm.world
        out.append({"workload": wl, "what": what, "chains_total": tot_chains, "chains_per_gpu": len(theta), "substeps": m,
                    "period": ds["period"], "params": M.d, "theta": kind, "steps": steps, "ms_per_step": ms / steps,
                    "value": tot_chains * steps / (ms * 1e-3), "unit": UNIT, "kernels": kdict(kms),
                    "executed_fp64_frac": fsum["whole_step_executed_frac"],
                    "survey_formula_frac": fsum["whole_step_survey_formula_frac"],
                    "fp64_pipe_pct_ncu": {k: v["fp64_pipe_pct_ncu"] for k, v in pk.items()} if ncu else None})
        del b
        M.close()
        torch.cuda.empty_cache()
    return out


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
    d = M.d
    mn, mx, init = M.df_params["min"], M.df_params["max"], M.df_params["init"]
    world = comm.world
    n = (args.chains + world - 1) // world if args.scaling == "strong" else args.chains
    total_chains = args.chains if args.scaling == "strong" else args.chains * world
    kind = theta_kind_of(args)
    # the configuration's batch is the same whatever N; a strong-scaled rank takes its slice of it
    whole = batch_theta(kind, args.workload, args.chains, mn, mx, init)
    theta = whole[comm.rank * n:(comm.rank + 1) * n] if args.scaling == "strong" else whole
    n = len(theta)
    ncu_all = load_ncu_constants()
    ncu = ncu_all.get(f"{args.workload}/m{args.substeps}/{kind}")
    peak_tf = M.fp64_peak_tflops()
    stream = torch.cuda.Stream(device=dev)   # the kernels are launched on THIS stream; so are the timing events
    stream.wait_stream(torch.cuda.current_stream(dev))

    clocks = ClockSampler(dev)
    clocks.start()
    time.sleep(0.3)   # let nvidia-smi start streaming before the timed region
    # ---- device-resident arm (headline): theta SoA already in HBM when the timed region starts
    head, hb, kms, launches, per_kernel, fsum = device_leg(M, ds, theta, args, comm, dev, stream, args.steps, args.warmup,
                                                           peak_tf, ncu, total_chains)
    value, step_ms = head["value"], head["ms_per_step"]
    # ---- the same step for >= 1 s (the contract's K steps last ~60 ms at K = 20)
    sus_steps = int(max(args.steps, np.ceil(1100.0 / step_ms)))
    ms_s, _ = timed(hb, sus_steps, 1, comm, dev)
    sustained = {"steps": sus_steps, "ms_per_step": ms_s / sus_steps, "value": total_chains * sus_steps / (ms_s * 1e-3), "unit": UNIT}
    ref_logl = hb.logl.cpu().numpy().copy()

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
    e2e_block_value = total_chains * args.steps / e2e_block_s
    assert np.allclose(ll_host[0].numpy(), ref_logl, rtol=0, atol=0, equal_nan=True)
    ll_host[0].zero_(); ll_host[1].zero_()
    pipelined(max(2, args.warmup))
    comm.barrier()
    t0 = time.perf_counter()
    pipelined(args.steps)
    e2e_s = comm.max_over_ranks(time.perf_counter() - t0)
    e2e_value = total_chains * args.steps / e2e_s
    for b in range(2):
        assert np.allclose(ll_host[b].numpy(), ref_logl, rtol=0, atol=0, equal_nan=True)
    clocks.stop()

    # ---- the narrow cloud of round 1 (warps in lock-step) beside the headline
    narrow = None
    if kind == "posterior":
        wn = batch_theta("narrow", args.workload, args.chains, mn, mx, init)
        tn = wn[comm.rank * n:(comm.rank + 1) * n] if args.scaling == "strong" else wn
        narrow, _, _, _, _, _ = device_leg(M, ds, tn, args, comm, dev, stream, args.steps, args.warmup, peak_tf, None, total_chains)
        narrow["theta"] = THETA_LABEL["narrow"]
    # ---- weak-scaling figure of a strong run: every GPU evaluates the whole configuration's batch
    weak = None
    if args.scaling == "strong" and world > 1:
        weak, _, _, _, _, _ = device_leg(M, ds, whole, args, comm, dev, stream, args.steps, args.warmup, peak_tf, None,
                                         args.chains * world)
        weak["chains_per_gpu"] = args.chains

    dom = KNAMES[int(np.argmax(kms))]
    r = per_kernel[dom]
    # frac: executed flop of the capture when there is one for this workload; otherwise the survey formula, capped
    # so that no fraction above 1 is ever printed as a roofline fraction
    frac_src = "executed FP64 flop per chain of profiles/ncu_constants.json (ncu SASS counts of this workload) x chains / launch time"
    ach, frac = r["achieved"], r["frac"]
    if ach is None:
        ach, frac = r["survey_formula_tflops"], r["survey_formula_frac"]
        frac_src = "SURVEY 8d formula with the reductions applied (no ncu capture of this workload)"
        if frac and frac > 1:
            ach, frac = None, None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": make_config(args, ds["period"], n, world, d, kind, _ws_mb(ds, n)),
        "clocks": clocks.summary(),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * d * 8), "d2h_bytes_per_step": int(n * 20),
                "bytes_are": "per GPU",
                "api": "epi_eval_submit/epi_eval_wait, two batches in flight (input copy of step k+1 overlaps the kernels of step k)",
                "blocking_call_value": e2e_block_value, "blocking_call_api": "epi_eval, one blocking call per step"},
        "gpu_launches": int(launches),
        "kernels": head["kernels"],
        "sustained": sustained,
        "roofline": {"bound": "fp64", "kernel": r["kernel"], "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": frac,
                     "traffic": r["traffic"], "frac_source": frac_src, "fp64_pipe_pct_ncu": r["fp64_pipe_pct_ncu"],
                     "survey_formula_frac": r["survey_formula_frac"],
                     "peak_source": "measured live: DFMA-chain microbenchmark epi_fp64_peak (MEASURED_PEAKS.json has no FP64 entry; "
                                    "nominal 37.2 TFLOP/s = 148 SM x 64 DFMA/clk x 2 x 1.965 GHz)",
                     "hbm": r["hbm"],
                     "traffic_source": ("dram__bytes_read.sum + dram__bytes_write.sum per launch, ncu --set full, profiles/"
                                        + str((ncu or {}).get("source"))) if r["traffic"] else None},
        "roofline_all": per_kernel,
        "f_alg": fsum,
        "narrow_cloud": narrow,
        "weak": weak,
    }
    if not args.no_secondary:
        line["secondary"] = secondary_legs(args, comm, dev, stream, peak_tf, ncu_all)
    if comm.rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(ds, theta, args.substeps, workload=args.workload)
    if not args.no_sampler:
        line["sampler"] = sampler_leg(M, comm, args, n)
    if comm.rank == 0:
        emit(line)
    M.close()
    comm.close()


def _n_terms(M):
    # rows of the likelihood term table (a-7): age-2q with n obs days: 2(n+2) + (n+10) + 1 + 2n
    ds = M.data
    if ds["flavour"] in ("age2q", "age2x"):
        n = ds["n_dcasei"]
        return 2 * (n + 2) + (ds["n_dhospi"] + 10) + 1 + 2 * ds["n_dmorti"] + (16 if ds["flavour"] == "age2x" else 0)
    if ds["flavour"] == "age2i":
        return 2 * (ds["n_dcasei"] + 2) + ds["n_dhospi"] + 2 * ds["n_dmorti"]
    return ds["n_dhospi"] + ds["n_dmorti"] + 1


def _ws_mb(ds, n):
    P = ds["period"]
    if ds["flavour"] in ("age2q", "age2x"):
        return n * ((P + 60) * 2 * 8 + 6 * 64 * 8) / 1e6
    if ds["flavour"] == "age2i":
        return n * (2 * P * 2 * 8 + 2 * 80 * 4 * 8) / 1e6
    return n * (2 * P * 8 + 2 * 64 * 8) / 1e6


def k2_roofline(M, chains_list=(65536, 1 << 22)):
    """the fused accept/propose kernel alone (epi_k2_bench): device time per launch, algorithmic bytes, GB/s against
    the measured HBM copy bandwidth.  65,536 chains x 45 parameters live in L2; 2^22 chains do not."""
    import ctypes as C
    from epi_mcmc_b200 import _capi
    hbm, src = hbm_peak()
    out = []
    for nc in chains_list:
        us, by, ar = C.c_double(), C.c_double(), C.c_double()
        _capi.check(M._lib.epi_k2_bench(M._h, _capi.SAMPLER_DEMC, int(nc), 50, C.byref(us), C.byref(by), C.byref(ar)), M._h)
        gbs = by.value / (us.value * 1e-6) / 1e9
        out.append({"chains": int(nc), "kernel": "k_accept_propose (DE-MC, subspace updates)", "us_per_iter": us.value,
                    "algorithmic_bytes": by.value, "achieved_GBs": gbs, "peak_GBs": hbm, "frac": gbs / hbm,
                    "accept_rate": ar.value, "population_MB": nc * M.d * 8 / 1e6, "peak_source": src})
    return out


def sampler_leg(M, comm, args, n):
    """DE-MC with the fused propose/accept kernel: chain state stays on the GPU; the population is all-gathered over
    NCCL every 10 iterations (side stream, double-buffered).  The population runs as B independent replicate
    populations (EPI_SAMPLER_FLAG_REPLICATES; a replicate spans all ranks): ESS of the whole run for a posterior mean
    = B var_posterior / var(replicate means), which needs no assumption about the dependence of the chains inside a
    population.  Reported for the whole recorded phase and for its first half (a stationary run gives the same
    ESS/s for both), with split R-hat over one chain per replicate."""
    import torch
    from epi_mcmc_b200.sampler import Sampler
    B = 64 if n % 64 == 0 and n // 64 >= 8 else (8 if n % 8 == 0 else 1)
    S = Sampler(M, n, kind="demc", seed=42, chain_id0=comm.rank * n, subspace=not args.sampler_fulldim, rungs=B,
                replicates=B > 1)
    comm.attach(S)
    burn, iters = args.sampler_burnin, args.sampler_iters
    thin = 10
    S.adapt(True)
    torch.cuda.synchronize()
    comm.barrier()
    tb = time.perf_counter()
    for _ in range(burn // 10):
        S.run(10)
        comm.exchange(S)
    S.adapt(False)
    comm.exchange(S, wait=True)   # the archive of the recorded phase: the population at the end of burn-in
    S.record(thin, iters // thin)
    torch.cuda.synchronize()
    comm.barrier()
    burn_s = comm.max_over_ranks(time.perf_counter() - tb)
    ev0 = S.stats()["evals"]
    acc0 = S.stats()["accepted"]
    t0 = time.perf_counter()
    for _ in range(iters // 10):
        S.run(10)
    torch.cuda.synchronize()
    dt = comm.max_over_ranks(time.perf_counter() - t0)
    st = S.stats()
    # the evaluation kernels alone on the sampler's OWN population (its chains have drifted apart further than the
    # Gaussian summary the headline cloud is drawn from): what an iteration costs beyond that is K2 + sequencing
    th_ptr, ll_ptr, lp_ptr, ld = S.state_device()
    scr = [torch.empty(n, dtype=torch.float64, device=f"cuda:{M.device}") for _ in range(2)]
    scr_st = torch.empty(n, dtype=torch.int32, device=f"cuda:{M.device}")
    pstream = torch.cuda.Stream(device=M.device)
    def pop_step():
        M.eval_device_range(th_ptr, ld, n, 0, n, scr[0].data_ptr(), scr[1].data_ptr(), scr_st.data_ptr(), pstream.cuda_stream)
    for _ in range(3):
        pop_step()
    torch.cuda.synchronize()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record(pstream)
    for _ in range(20):
        pop_step()
    p1.record(pstream)
    torch.cuda.synchronize()
    pop_ms = comm.max_over_ranks(p0.elapsed_time(p1)) / 20
    keys = [M.fit_paramnames.index(k) for k in M.fitkeyparamnames if k in M.fit_paramnames]
    full = S.replicate_stats(keys, comm) if B > 1 else None
    half = S.replicate_stats(keys, comm, last=(iters // thin) // 2) if B > 1 else None
    evals = comm.world * (st["evals"] - ev0)
    out = {"kind": "DE-MC (scale mixture, subspace updates); burn-in: population snapshot all-gathered every 10 iterations "
                   "(side stream, double-buffered) and proposal shape re-measured; recorded phase: both frozen",
           "chains_per_gpu": n, "replicate_populations": B * 1, "chains_per_replicate": (n // B) * comm.world,
           "burnin_iterations": burn, "burnin_s": burn_s, "burnin_evals_per_s": comm.world * n * burn / burn_s,
           "sampling_iterations": iters, "sampling_s": dt, "thin": thin,
           "evals_per_s": evals / dt, "accept_rate": (st["accepted"] - acc0) / max(1, (st["evals"] - ev0)),
           "ms_per_iteration": 1e3 * dt / iters, "eval_kernels_on_this_population_ms": pop_ms,
           "evals_per_s_vs_bare_eval_of_this_population": (1e3 * dt / iters and pop_ms / (1e3 * dt / iters))}
    if full is not None:
        e_full, e_half = float(np.median(full["ess"])), float(np.median(half["ess"]))
        out.update({"ess_median": e_full, "ess_median_first_half": e_half,
                    "ess_per_s_median": e_full / dt, "ess_per_s_min": float(full["ess"].min()) / dt,
                    "ess_per_s_median_first_half": e_half / (dt / 2),
                    "ess_growth_per_s": (e_full - e_half) / (dt / 2),
                    "ess_per_eval_median": e_full / max(1, evals),
                    "ess_note": "ESS of the posterior-mean estimator over ALL chains and recorded iterations.  When the recorded "
                                "window is shorter than a chain's integrated autocorrelation time, the ESS is carried by the "
                                "spread of the population at the start of the window and barely grows with it (ess_median vs "
                                "ess_median_first_half, split R-hat well above 1): ESS / window is then NOT an asymptotic "
                                "rate — ess_growth_per_s is the rate at which the window adds information",
                    "split_rhat_max": float(full["rhat"].max()), "split_rhat_median": float(np.median(full["rhat"])),
                    "ess_method": f"between-replicate variance of the posterior-mean estimator over {full['n_replicates']} independent "
                                  f"replicate populations (relative sd of the estimate ~ sqrt(2 / {full['n_replicates'] - 1})), "
                                  f"{len(keys)} key parameters; split R-hat over one chain per replicate"})
    # ---- the drjacoby-style kernel on the same target, measured the same way, for the same number of evaluations per
    # chain as the recorded DE-MC phase, started from the DE-MC population (like-for-like with the reference's sampler)
    if B > 1:
        start = S.get()[0]
        nfree = int((np.asarray(M.df_params["max"]) > np.asarray(M.df_params["min"])).sum())
        sweeps = max(8, iters // nfree)
        D = Sampler(M, n, kind="drj", seed=43, chain_id0=comm.rank * n, rungs=B, replicates=True, theta0=start)
        D.adapt(True)
        D.run(20)
        D.adapt(False)
        D.record(1, sweeps)
        e0, a0 = D.stats()["evals"], D.stats()["accepted"]
        torch.cuda.synchronize()
        comm.barrier()
        t0 = time.perf_counter()
        D.run(sweeps)
        torch.cuda.synchronize()
        ddt = comm.max_over_ranks(time.perf_counter() - t0)
        dst = D.stats()
        dfull, dhalf = D.replicate_stats(keys, comm), D.replicate_stats(keys, comm, last=sweeps // 2)
        devals = comm.world * (dst["evals"] - e0)
        df_, dh_ = float(np.median(dfull["ess"])), float(np.median(dhalf["ess"]))
        out["drj"] = {"kind": "drjacoby's own update (one parameter at a time on the logit-reparameterised box, Robbins-Monro "
                              "bandwidths frozen after 20 sweeps), started from the DE-MC population above",
                      "sweeps": sweeps, "evals_per_chain": (dst["evals"] - e0) // n, "sampling_s": ddt, "evals_per_s": devals / ddt,
                      "accept_rate": (dst["accepted"] - a0) / max(1, dst["evals"] - e0),
                      "ess_median": df_, "ess_median_first_half": dh_, "ess_per_s_median": df_ / ddt,
                      "ess_growth_per_s": (df_ - dh_) / (ddt / 2), "ess_per_eval_median": df_ / max(1, devals),
                      "split_rhat_median": float(np.median(dfull["rhat"])), "split_rhat_max": float(dfull["rhat"].max())}
    if comm.rank == 0:
        out["k2"] = k2_roofline(M)
    return out


if __name__ == "__main__":
    main()
