#!/usr/bin/env python
"""bench.py -- I(q) pattern evaluations per second and batched LM fits per second on B200, beside the reference's
CPU path (BASELINE.json: "I(q) pattern evals/sec & batched LM fits/sec at 1/2/4/8 B200 vs host CPU; % roofline").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...        # the CPU arm: the unmodified reference on the host cores

Headline workload (default): BASELINE.json configs[1] = "cfg2": polydisperse spheres, 50-pt r_normal quadrature,
hard-sphere S(q), 4096-pt q grid, 10k parameter sets per GPU.  A *step* is `passes_per_step` passes of the hot path
over that batch (one launch each, so that a step is ~50 ms of GPU work and a 20-step run ~1 s).  Under torchrun
every rank runs the same per-GPU batch (weak scaling); rank 0 prints ONE JSON line.

value        units/s with inputs resident in HBM (CUDA events around K steps, max over ranks)
e2e          the same through the public host-buffer API: pinned host inputs in, results out to pinned host
             memory, copies inside the timed region
roofline     dominant kernel against the FP64 (DFMA) pipe peak measured live by a probe kernel (and against the
             nominal 148 SM x 64 DFMA/clk x 1.965 GHz), with achieved HBM GB/s against MEASURED_PEAKS.json beside it
cpu_baseline the reference's CPU path on the host cores, bounded sample (oracle/cpu_arm.py: the unmodified
             reference from oracle/_ref when it travelled with the snapshot, else the pinned numpy port)
secondary    (N = 1, default workload) the other BASELINE configs at full batch, each with its own roofline / e2e /
             cpu_baseline: cfg3 (FCC diffraction), cfg4 (residual + FD Jacobian), cfg5 (a mixed batch of sphere and
             crystal LM fits with success fractions, beside batched Nelder-Mead on the same systems)
             (N > 1) cfg5_gathered: every rank fits its shard and the (parameters, chi2) rows meet in one NCCL
             all_gather inside the timed region
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

WL = 0.8265617
FP64_NOMINAL_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12      # 148 SMs x 64 DFMA / clk x 2 flop x 1.965 GHz = 37.2
SUCCESS_EPS = 0.01      # a fit succeeded when chi2_final <= chi2(truth) * (1 + eps)   
JITTER = 0.10           # start = truth * (1 + 0.1 U(-1, 1))                          
LM_ITERS = 30           # iteration cap                                              

# Algorithmic FP64 work per unit, as stated in DESIGN.md section 4 (FMA = 2, mul/add = 1, div = 10,
# sincos pair = 40, exp = 30, Re w = the series actually needed):
#   cfg2: per (set, q): n_r * 13 (rotation 4 ops = 6 flop, x 1, t 2, t^2 1, acc 2 ... = 13) + 151
FLOPS_PER_PAIR_RNORMAL = 13.0
FLOPS_PER_Q_OVERHEAD_CFG2 = 151.0        # PY brackets ~100, two divisions 20, three span-to-span rotations 18, epilogue ~13
SURVEY_FLOPS_PER_SETQ_CFG2 = 2630.0      # SURVEY.md 8d convention (sincos per (q,r) pair)
#   crystalline: one far-branch Voigt evaluation = reciprocal (3 FMA) + set-up (mul, add, FMA) + Clenshaw chain of the
#   asymptotic series (9 FMA at k <= 5) + finish (2 mul, FMA) + argument and accumulation (1 mul, 2 FMA) = 21 FP64
#   instructions, 37 flop (23 instructions / 41 flop with the forward recurrence used until r02h)
VOIGT_FLOPS = 37.0
#   near-branch Voigt evaluation through the pattern's table (|z| < 8): interval and argument 4, u = 2t^2 - 1 and two
#   Horner chains of 6 FMA, t O + E, accumulation: 17 FP64 instructions, 31 flop
VOIGT_TABLE_FLOPS = 31.0
FAR_NODES = 15.0         # Chebyshev nodes of the far-field interpolant per run of 128 q-points (csrc/system_eval.cuh, CHEB_N)

WORKLOADS = {
    "cfg2": "cfg2: polydisperse spheres (r_normal, 50-pt quadrature) + hard_spheres S(q), 4096-pt q, 10k parameter sets per GPU",
    "cfg3": "cfg3: crystalline FCC Fm-3m, spherical form, voigt, q_max=1.0, 8192-pt q, table build + profile sum",
    "cfg4": "cfg4: FCC crystal + guinier_porod + flat noise, 8192-pt q, residual + FD Jacobian (8 free parameters)",
    "cfg5": "cfg5: batched Levenberg-Marquardt fit of a mixed batch (half cfg2-like sphere systems on 4096 q, half "
            "cfg4-like crystal systems on 8192 q), I_obs = I(truth)(1 + 0.02 randn), start = truth(1 + 0.1 U(-1,1)), 30 iterations",
}
N_Q = {"cfg2": 4096, "cfg3": 8192, "cfg4": 8192, "cfg5": 8192}
UNITS = {"cfg2": ("I(q) pattern evals/sec", "patterns/s"), "cfg3": ("I(q) pattern evals/sec", "patterns/s"),
         "cfg4": ("I(q) pattern evals/sec (residual + Jacobian, reference-equivalent evaluations)", "patterns/s"),
         "cfg5": ("batched LM fits/sec", "fits/s")}
PASSES = {"cfg2": 36, "cfg3": 8, "cfg4": 2, "cfg5": 1}


def config_of(workload, batch):
    """The `config` object of the JSON line: static description of the workload, identical on both arms."""
    return {"workload": WORKLOADS[workload], "n_q": N_Q[workload], "sets_per_gpu_per_step": batch,
            "passes_per_step": PASSES[workload],
            "evaluations_per_set": 9.0 if workload == "cfg4" else 1.0,
            "l2": "every pass writes or reads %.0f MB of patterns (L2 is 126 MB) and rotates over 4 parameter matrices"
                  % (8e-6 * batch * N_Q[workload])}


# --------------------------------------------------------------------------------------
# CPU arm (imports oracle/ only -- never the product package)
# --------------------------------------------------------------------------------------
def run_reference_arm(args):
    from oracle import cpu_arm
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    cores = cpu_arm.host_cores()
    per_core = {"cfg2": 128, "cfg3": 1, "cfg4": 1, "cfg5": 1}[wl]      # per step and core: ~1 s of work, so the pool's hand-off does not show
    metric, unit = UNITS[wl]
    if wl == "cfg5":
        line = _reference_fit_line(args, cpu_arm, cores)
    else:
        arm = cpu_arm.CpuArm(wl, cores)
        for w in range(args.warmup):
            arm.step(per_core, w * cores * per_core)
        t_tot, n_tot = 0.0, 0
        for k in range(args.steps):
            n, dt = arm.step(per_core, (args.warmup + k) * cores * per_core)
            t_tot += dt
            n_tot += n
        arm.close()
        val = n_tot / t_tot
        line = {"value": val, "ms_per_step": 1e3 * t_tot / max(1, args.steps),
                "cpu_baseline": {"value": val, "unit": unit, "cores": cores, "kind": arm.kind,
                                 "sample": "%d steps x %d parameter sets (%d per core) of the same synthetic workload, %s, one "
                                           "process per core" % (args.steps, cores * per_core, per_core, cpu_arm.describe(arm.kind))}}
    line.update({"impl": "reference", "metric": metric, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
                 "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                 "data": "synthetic", "config": config_of(wl, args.batch),
                 "e2e": {"value": line["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))
    return 0


def _reference_fit_line(args, cpu_arm, cores):
    """cfg5 on the host cores: the sphere half fitted the reference's way (Nelder-Mead on its scalar objective, one
    system per core); a crystal fit costs ~10^3 evaluations of ~1 s each, so its rate is the measured crystal
    pattern rate divided by the evaluations the sphere fits needed (stated in `sample`)."""
    arm = cpu_arm.CpuArm("cfg5s", cores)
    lay, q = arm_layout(cpu_arm, "cfg5s")
    truth = cpu_arm.workload_params("cfg5s", lay, cores)
    free = [lay.slot(k) for k in cpu_arm.FREE_NAMES["cfg2"]]
    start, iobs = _cpu_fit_inputs(cpu_arm, "cfg5s", truth, free)
    t0 = time.perf_counter()
    rate_s, recs = arm.fits(start, iobs, free)
    dt = time.perf_counter() - t0
    arm.close()
    evals = float(np.median([r[1] for r in recs]))
    arm_x = cpu_arm.CpuArm("cfg4", cores)
    n, dtx = arm_x.step(1)
    arm_x.close()
    rate_x = (n / dtx) / evals
    mixed = 2.0 / (1.0 / rate_s + 1.0 / rate_x)
    return {"value": mixed, "ms_per_step": 1e3 * (dt + dtx),
            "cpu_baseline": {"value": mixed, "unit": "fits/s", "cores": cores, "kind": arm.kind,
                             "sample": "%d sphere systems fitted, one per core (scipy Nelder-Mead, xatol = fatol = 1e-4, on %s; "
                                       "median %d evaluations, %.2f fits/s); crystal half EXTRAPOLATED: %d cfg4 patterns timed "
                                       "(%.2f patterns/s) / the same evaluation count = %.4f fits/s; 50/50 mix = harmonic mean"
                                       % (len(recs), cpu_arm.describe(arm.kind), evals, rate_s, n, n / dtx, rate_x)}}


def arm_layout(cpu_arm, workload):
    _, lay, q = cpu_arm.load_workload(workload)
    return lay, q


def _cpu_fit_inputs(cpu_arm, workload, truth, free, seed=4242):
    """start rows and observed patterns for CPU fits of `truth` rows (the GPU arm's recipe, numpy generator)."""
    be = cpu_arm.Backend(workload)
    rs = np.random.RandomState(seed)
    iobs = np.stack([be.intensity(row) * (1 + 0.02 * rs.randn(len(be.q))) for row in truth])
    start = truth.copy()
    start[:, free] *= 1 + JITTER * rs.uniform(-1, 1, (len(truth), len(free)))
    return start, iobs


# --------------------------------------------------------------------------------------
# clocks during the timed region (NVML, falls back to nvidia-smi)
# --------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.nv, self.h = None, None
        try:                         # NVML is initialised here, before the timed region starts
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _smi(self):
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
            a, b = out.strip().split(",")
            self.samples.append(int(a))
            self.max_mhz = int(b)
        except Exception:
            pass

    def run(self):
        if self.nv is None:
            while not self.stop_flag.is_set():
                self._smi()
            return
        nv, h = self.nv, self.h
        while not self.stop_flag.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    bits = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    bits = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in self.REASONS.items():
                    if bits & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def result(self):
        self.stop_flag.set()
        self.join(2.0)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------
def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


def make_workload(name, batch, seed=None):
    from helpers import cfg2_system, cfg2_params, cfg3_system, cfg3_params, cfg4_system, cfg4_params
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch as xb
    kw = {} if seed is None else {"seed": seed}
    if name in ("cfg2", "cfg5s"):
        s, q = cfg2_system(System), np.linspace(0.01, 0.5, 4096)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg2_params(lay, batch, **kw)
    if name == "cfg3":
        s, q = cfg3_system(System), np.linspace(0.01, 1.0, 8192)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg3_params(lay, batch, **kw)
    if name in ("cfg4", "cfg5x"):
        s, q = cfg4_system(System), np.linspace(0.01, 1.0, 8192)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg4_params(lay, batch, **kw)
    raise SystemExit("unknown workload " + name)


FREE_NAMES = {"cfg4": ("noise__I0", "x__I0", "x__a", "x__hwhm_g", "x__hwhm_l", "x__r", "gp__I0", "gp__rg"),
              "cfg5x": ("noise__I0", "x__I0", "x__a", "x__hwhm_g", "x__hwhm_l", "x__r", "gp__I0", "gp__rg"),
              "cfg5s": ("noise__I0", "p__I0", "p__r", "p__sigma", "p__r_hard", "p__v_fraction")}


class FitGroup(object):
    """The systems of one layout of a fit workload on one GPU: truth, observed patterns, start values."""

    def __init__(self, eng, name, B, seed, jitter):
        import torch
        self.eng, self.name, self.B = eng, name, B
        dev = eng.device
        self.s, self.q, self.lay, self.P = make_workload(name, B, seed)
        self.n_q = len(self.q)
        self.d_q = eng.to_device(self.q)
        self.free = [self.lay.slot(k) for k in FREE_NAMES[name]]
        self.truth = eng.to_device(self.P)
        g = torch.Generator(dev).manual_seed(seed + 5)
        self.Iobs = eng.intensity(self.lay, self.d_q, self.truth).clone()
        self.Iobs *= 1 + 0.02 * torch.randn(self.Iobs.shape, dtype=torch.float64, device=dev, generator=g)
        self.obj = eng.objective(True, True, (0.0, float("inf")))
        jit = 1 + jitter * (2 * torch.rand(self.truth.shape, dtype=torch.float64, device=dev, generator=g) - 1)
        mask = torch.zeros(self.lay.n_params, dtype=torch.bool, device=dev)
        mask[self.free] = True
        start = torch.where(mask[None, :], self.truth * jit, self.truth)
        lo, hi = self.lay.bounds(list(range(self.lay.n_params)))      # lmfit clips a start value to its bounds
        self.start = torch.minimum(torch.maximum(start, eng.to_device(np.where(np.isfinite(lo), lo, -1e300))[None, :]),
                                   eng.to_device(np.where(np.isfinite(hi), hi, 1e300))[None, :]).contiguous()
        self.chi2_truth = eng.residual(self.lay, self.obj, self.d_q, self.truth, self.Iobs).clone()

    def group(self, rows=None):
        sl = slice(0, rows)
        return dict(layout=self.lay, q=self.d_q, Iobs=self.Iobs[sl], params0=self.start[sl], free_slots=self.free)

    def quality(self, res, rows=None):
        """success fraction (chi2 <= chi2(truth)(1+eps)) and chi2 ratios of a FitResult over the first `rows` systems"""
        import torch
        ct = self.chi2_truth[:rows] if rows else self.chi2_truth
        r = res.chi2 / ct
        ok = torch.isfinite(r) & (r <= 1.0 + SUCCESS_EPS)
        return {"systems": int(r.numel()), "success_frac": float(ok.double().mean().item()),
                "median_chi2_over_truth": float(torch.nan_to_num(r, nan=float("inf")).median().item()),
                "median_chi2_ratio": float(torch.nan_to_num(res.chi2 / res.chi2_init, nan=float("inf")).median().item())}


class Work(object):
    """One workload instance on one GPU: device buffers + a step() closure."""

    def __init__(self, eng, name, B, seed, lm_iters=LM_ITERS, jitter=JITTER):
        import torch
        from xrsdkit_b200 import lowering
        self.eng, self.name, self.B = eng, name, B
        self.passes = PASSES[name]
        self.metric, self.unit = UNITS[name]
        self.evals_per_unit = 9.0 if name == "cfg4" else 1.0
        self.lm_iters = lm_iters
        dev = eng.device
        if name == "cfg5":
            # a mixed batch: B/2 sphere systems and B/2 crystal systems, fitted side by side (batch.fit_mixed)
            self.groups = [FitGroup(eng, "cfg5s", B // 2, seed, jitter), FitGroup(eng, "cfg5x", B - B // 2, seed + 1, jitter)]
            self.n_q = self.groups[1].n_q
            self.res = None
            return
        if name == "cfg4":
            # residual + Jacobian at start points 3 % off the truth (a typical LM iterate); I_obs as in cfg5
            g = FitGroup(eng, "cfg4", B, seed, 0.03)
            self.fit = g
            self.s, self.q, self.lay, self.P, self.n_q, self.d_q = g.s, g.q, g.lay, g.P, g.n_q, g.d_q
            self.free, self.Iobs, self.start, self.obj = g.free, g.Iobs, g.start, g.obj
            self.jout = None
            self.n_var = eng.jacobian_variants(self.lay, self.free)
            self.max_cells = lowering.max_cells(self.lay, self.start.cpu().numpy())
            self.tables = [eng.build_tables(self.lay, self.start, B, max_cells=self.max_cells) for _ in range(4)]
            return
        self.s, self.q, self.lay, self.P = make_workload(name, B)
        self.n_q = len(self.q)
        self.d_q = eng.to_device(self.q)
        rs = np.random.RandomState(seed)
        # 4 distinct parameter matrices and 2 output buffers are rotated: one pass' output alone
        # (8 B x B x n_q) exceeds the 126 MB L2 for the full-size batches
        self.d_params = [eng.to_device(self.P[rs.permutation(B)]) for _ in range(4)]
        self.max_cells = lowering.max_cells(self.lay, self.P) if self.lay.n_xtal else None
        self.tables = [eng.build_tables(self.lay, p, B, params_host=self.P) if self.lay.n_xtal else None for p in self.d_params]
        self.outs = [torch.empty((B, self.n_q), dtype=torch.float64, device=dev) for _ in range(2)]

    def one_pass(self, i):
        eng, B = self.eng, self.B
        if self.name == "cfg5":         # a whole batched Levenberg-Marquardt fit of the mixed batch
            from xrsdkit_b200 import batch as xb
            self.res = xb.fit_mixed([g.group() for g in self.groups], eng=eng, max_iter=self.lm_iters)
            return
        lay = self.lay
        if self.name == "cfg4":      # table build + residual + forward-difference Jacobian (9 reference evaluations per system, fused)
            k = i % 4
            self.tables[k].store.reset_shared()       # the distinct tables are rebuilt in every pass: part of the evaluation
            t = eng.build_tables(lay, self.start, B, max_cells=self.max_cells, check=False, reuse=self.tables[k])
            self.jout = eng.jacobian(lay, self.obj, self.d_q, self.start, self.Iobs, None, self.n_q, self.free, t, 1e-6, out=self.jout)
        else:
            k = i % 4
            t = self.tables[k]
            if self.name == "cfg3":  # the reflection tables are rebuilt in every pass: part of the pattern evaluation
                t.store.reset_shared()
                t = eng.build_tables(lay, self.d_params[k], B, max_cells=self.max_cells, check=False, reuse=t)
            eng.intensity(lay, self.d_q, self.d_params[k], tables=t, out=self.outs[i % 2])

    def step(self, i):
        for p in range(self.passes):
            self.one_pass(i * self.passes + p)

    def timed(self, steps, warmup, barrier=None):
        import torch
        for i in range(warmup):
            self.step(i)
        (barrier or torch.cuda.synchronize)()
        n0 = self.eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            self.step(i)
        e1.record()
        (barrier or torch.cuda.synchronize)()
        return e0.elapsed_time(e1), self.eng.launches - n0

    def units_per_step(self):
        return self.B * self.passes * self.evals_per_unit

    # ---- e2e: the same work through host buffers (pinned), copies inside the timed region ----
    def make_e2e(self):
        import torch
        from xrsdkit_b200 import batch as xb
        eng, B = self.eng, self.B
        if self.name in ("cfg2", "cfg3"):
            lay, n_q = self.lay, self.n_q
            p_host = torch.as_tensor(self.P).pin_memory()
            q_host = torch.as_tensor(self.q).pin_memory()
            I_host = torch.empty((B, n_q), dtype=torch.float64).pin_memory()

            def fn():
                for _ in range(self.passes):
                    xb.compute_intensity_pinned(lay, q_host, p_host, I_host, eng=eng)
            return (fn, self.passes * int(self.P.nbytes + self.q.nbytes), self.passes * int(8 * B * n_q),
                    "xrsdkit_b200.batch.compute_intensity_pinned (pinned host buffers, chunked copy/compute overlap)",
                    lambda: float(I_host[:: max(1, B // 7)].sum().item()))
        if self.name == "cfg4":
            Iobs_host, start_host = self.Iobs.cpu().pin_memory(), self.start.cpu().pin_memory()
            q_host = torch.as_tensor(self.q).pin_memory()
            nf = len(self.free)
            out_host = torch.empty((B, nf * nf + nf + 1), dtype=torch.float64).pin_memory()

            def fn():
                for p in range(self.passes):
                    self.Iobs.copy_(Iobs_host, non_blocking=True)
                    self.start.copy_(start_host, non_blocking=True)
                    self.d_q.copy_(q_host, non_blocking=True)
                    self.one_pass(p)
                    JTJ, JTr, chi2 = self.jout
                    out_host.copy_(torch.cat([JTJ.reshape(B, -1), JTr, chi2[:, None]], dim=1), non_blocking=True)
                torch.cuda.synchronize()
            return (fn, self.passes * int(Iobs_host.numel() * 8 + start_host.numel() * 8 + self.q.nbytes),
                    self.passes * int(out_host.numel() * 8),
                    "engine.build_tables + engine.jacobian with pinned host I_obs / start parameters in, normal equations out",
                    lambda: float(out_host[:: max(1, B // 7)].nansum().item()))
        # cfg5: pinned I_obs and start parameters of both groups in, fitted parameters + chi2 out
        hosts = [(g.Iobs.cpu().pin_memory(), g.start.cpu().pin_memory(),
                  torch.empty((g.B, g.lay.n_params + 1), dtype=torch.float64).pin_memory()) for g in self.groups]

        def fn():
            for g, (Ih, sh, _) in zip(self.groups, hosts):
                g.Iobs.copy_(Ih, non_blocking=True)
                g.start.copy_(sh, non_blocking=True)
            self.one_pass(0)
            for r, (_, _, oh) in zip(self.res, hosts):
                oh.copy_(torch.cat([r.params, r.chi2[:, None]], dim=1), non_blocking=True)
            torch.cuda.synchronize()
        return (fn, sum(int(Ih.numel() * 8 + sh.numel() * 8) for Ih, sh, _ in hosts), sum(int(oh.numel() * 8) for _, _, oh in hosts),
                "xrsdkit_b200.batch.fit_mixed with pinned host I_obs / start parameters in, fitted parameters and chi2 out",
                lambda: float(sum(oh[:: max(1, oh.shape[0] // 7)].nansum().item() for _, _, oh in hosts)))

    # ---- algorithmic FP64 work ----
    def xtal_flops_per_q(self, lay, P, q, tables):
        """flop per (pattern, q) of ONE crystalline evaluation as built (far-field interpolation: per q only the shells
        within dn = max(4H, H + 12 core widths) of its 128-point run are evaluated directly, the rest through 15 Chebyshev
        nodes per run), and by the SURVEY 8d convention (205 per shell and q)."""
        n_q = len(q)
        ns = float(np.mean([float(t.n_shell.float().mean().item()) for t in tables]))
        H = 0.5 * 127 * (q[-1] - q[0]) / (n_q - 1)
        hg, hl = P[:, lay.slot("x__hwhm_g")], P[:, lay.slot("x__hwhm_l")]
        s2 = hg / np.sqrt(2 * np.log(2)) * np.sqrt(2)
        core = np.maximum(1.0, hl / s2) * s2
        dn = np.maximum(4 * H, H + 12 * core)
        n_near = np.minimum(ns, ns * 2 * dn / (q[-1] - q[0]))
        # the part of the near range within |z| < 8 (8 s2 of the peak) takes the table branch, the rest the far branch
        near_flops = float(np.mean(n_near * (VOIGT_FLOPS + (VOIGT_TABLE_FLOPS - VOIGT_FLOPS) * np.minimum(1.0, 8.0 * s2 / dn))))
        n_near = float(np.mean(n_near))
        # per q besides the shells: the interpolant (15 FMA + 2) = 2 FAR_NODES + 2 flop ... and ~150 for the rest
        return near_flops + FAR_NODES * VOIGT_FLOPS * (ns - n_near) / 128.0 + 2.0 * FAR_NODES + 150.0, 205.0 * ns + 150.0


def ncu_kernel_counts(workload, batch=10000):
    """Per-launch figures of the dominant kernel from the committed ncu capture of this round (profiles/): DRAM
    traffic and the FP64 thread-instruction counts (2 flop per DFMA, 1 per DADD / DMUL), for `batch` systems per launch
    (the capture records how many systems its launch held); None when the file is absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_ncu_counts.json")) as f:
            rec = json.load(f).get(workload)
        if rec is None:
            return None
        out = {k: v for k, v in rec.items() if k != "all"}
        # the capture may hold fewer systems per launch than the timed launch (the public Jacobian call cuts the batch
        # into <= 4 GB chunks): counts that scale with the systems are brought to the timed launch's size
        n_cap = rec.get("systems_per_launch")
        if n_cap and n_cap != batch:
            for k in ("dram_bytes", "fp64_flops"):
                if out.get(k) is not None:
                    out[k] = out[k] * batch / float(n_cap)
            out["scaled_from_systems"] = n_cap
        elif not n_cap and batch != 10000:
            return None
        return out
    except Exception:
        return None


def roofline_entry(kernel, k_ms, flops_unit, survey_unit, units, hbm_bytes, fp64_tflops, probe_ms, peaks, peaks_src, counts):
    achieved = flops_unit * units / (k_ms * 1e-3) / 1e12
    out = {
        "bound": "fp64", "achieved": achieved, "peak": fp64_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_tflops,
        "peak_nominal": FP64_NOMINAL_TFLOPS, "frac_nominal": achieved / FP64_NOMINAL_TFLOPS,
        "traffic": None, "kernel": kernel, "kernel_ms": k_ms, "flops_per_unit": flops_unit, "units_per_launch": units,
        "peak_source": "xrsd_probe_fp64_peak: DFMA-bound kernel timed in this run (%.3f ms; 8 independent DFMA chains per "
                       "thread, 1184 CTAs x 256 threads); MEASURED_PEAKS.json has no FP64 figure; peak_nominal = 148 SM x "
                       "64 DFMA/clk x 2 x 1.965 GHz" % probe_ms,
        "achieved_survey_convention": survey_unit * units / (k_ms * 1e-3) / 1e12,
        "hbm": {"achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peaks_src,
                "algorithmic_bytes_per_launch": hbm_bytes},
    }
    if counts and counts.get("executed"):
        out["executed"] = counts["executed"]
    if counts:
        out["traffic"] = counts.get("dram_bytes")
        if counts.get("fp64_flops"):
            # executed FP64 flops of one launch as ncu counted them (same launch shape), against this run's kernel time
            out["ncu"] = {"fp64_flops_per_launch": counts["fp64_flops"], "fp64_pipe_active_pct": counts.get("fp64_pipe_pct"),
                          "achieved_executed": counts["fp64_flops"] / (k_ms * 1e-3) / 1e12,
                          "frac_executed": counts["fp64_flops"] / (k_ms * 1e-3) / 1e12 / fp64_tflops,
                          "source": counts.get("source")}
    return out


def time_kernel(fn, reps=10):
    """Average duration of one launch of `fn` (CUDA events on the launching stream).  The median of three batches after
    a warm-up of 2 x reps launches: this runs after the host-buffer phase of the bench, during which the GPU idles between
    copies, and a single batch taken right then was seen 12 % slow."""
    import torch
    for i in range(2 * reps):
        fn(i)
    torch.cuda.synchronize()
    out = []
    for _ in range(3):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        for i in range(reps):
            fn(i)
        k1.record()
        torch.cuda.synchronize()
        out.append(k0.elapsed_time(k1) / reps)
    return sorted(out)[1]


def dominant_kernel(eng, W, fp64_tflops, probe_ms, peaks, peaks_src):
    """Roofline entry of the workload's dominant kernel, timed alone with CUDA events on the launching stream."""
    name, B = W.name, W.B
    if name in ("cfg2", "cfg3"):
        lay, n_q = W.lay, W.n_q
        k_ms = time_kernel(lambda i: eng.intensity(lay, W.d_q, W.d_params[i % 4], tables=W.tables[i % 4], out=W.outs[i % 2]))
        if name == "cfg2":
            fl, sv = FLOPS_PER_PAIR_RNORMAL * 50.5 + FLOPS_PER_Q_OVERHEAD_CFG2, SURVEY_FLOPS_PER_SETQ_CFG2
        else:
            fl, sv = W.xtal_flops_per_q(lay, W.P, W.q, W.tables)
        return roofline_entry("xrsd::intensity_kernel<2,false,%s>" % ("true" if lay.n_xtal else "false"), k_ms, fl, sv,
                              B * n_q, 8.0 * B * n_q, fp64_tflops, probe_ms, peaks, peaks_src, ncu_kernel_counts(name, B))
    if name == "cfg4":
        lay, n_q = W.lay, W.n_q
        key = ("bench", id(W))
        try:        # the variants kernel alone: a patterns-only launch into a resident workspace
            k_ms = time_kernel(lambda i: eng.jacobian(lay, W.obj, W.d_q, W.start, None, None, 0, W.free, W.tables[0], 1e-6,
                                                      resident=key, patterns_only=True), reps=5)
        finally:
            eng.release_resident(key)
        fl, sv = W.xtal_flops_per_q(lay, W.P, W.q, W.tables)
        n_x, n_gp = executed_evaluations(lay, W.free)
        # EXECUTED work per (system, q): the base pattern (crystal + Guinier-Porod) plus one evaluation of ITS population
        # per non-linear parameter (a, hwhm_g, hwhm_l, r -> crystal; rg -> Guinier-Porod); the I0 derivatives are analytic
        return roofline_entry("xrsd::jacobian_variants_kernel<false,true>", k_ms, n_x * fl + n_gp * 80.0, n_x * sv + n_gp * 80.0,
                              B * n_q, 8.0 * B * n_q * (W.n_var + lay.c.n_pops), fp64_tflops, probe_ms, peaks, peaks_src,
                              dict(ncu_kernel_counts(name, B) or {}, executed={"crystal_evaluations": n_x, "guinier_porod_evaluations": n_gp}))
    # cfg5: the dominant kernel of the fit is the same variants kernel, on the crystal group's layout
    g = W.groups[1]
    lay, n_q = g.lay, g.n_q
    tables = eng.build_tables(lay, g.start, g.B)
    key = ("bench", id(W))
    try:
        k_ms = time_kernel(lambda i: eng.jacobian(lay, g.obj, g.d_q, g.start, None, None, 0, g.free, tables, 1e-6,
                                                  resident=key, patterns_only=True), reps=5)
    finally:
        eng.release_resident(key)
    n_var = eng.jacobian_variants(lay, g.free)
    fl, sv = W.xtal_flops_per_q(lay, g.P, g.q, [tables])
    n_x, n_gp = executed_evaluations(lay, g.free)
    return roofline_entry("xrsd::jacobian_variants_kernel<false,true> (crystal group of the fit: %d systems)" % g.B, k_ms,
                          n_x * fl + n_gp * 80.0, n_x * sv + n_gp * 80.0, g.B * n_q, 8.0 * g.B * n_q * (n_var + lay.c.n_pops),
                          fp64_tflops, probe_ms, peaks, peaks_src,
                          dict(ncu_kernel_counts("cfg5", g.B) or {}, executed={"crystal_evaluations": n_x, "guinier_porod_evaluations": n_gp}))


def executed_evaluations(lay, free):
    """(crystalline, Guinier-Porod) population evaluations per system of one variants launch: the base pattern evaluates
    every population; each non-linear free parameter re-evaluates the one population it belongs to."""
    n_x, n_gp = 1, 1
    for slot in free:
        name = lay.names[slot]
        if name.endswith("__I0"):
            continue
        if name.startswith("x__"):
            n_x += 1
        elif name.startswith("gp__"):
            n_gp += 1
    return n_x, n_gp


def fit_quality(eng, W, nm_rows, with_nm=True):
    """Quality of the LM fits of a cfg5 Work (all systems) and, on the first `nm_rows` systems of each group, beside
    the reference's optimiser (batched Nelder-Mead, xrsdkit_b200.fitting.nm_fit_batch) on the same systems."""
    import torch
    from xrsdkit_b200 import fitting
    out = {"success_eps": SUCCESS_EPS, "start_jitter": JITTER, "lm_iterations": W.lm_iters}
    for g, r in zip(W.groups, W.res):
        kind = "sphere" if g.name == "cfg5s" else "crystal"
        out[kind] = {"lm": g.quality(r)}
        out[kind]["lm"]["converged_frac"] = float(r.converged.double().mean().item())
        if with_nm:
            rows = min(nm_rows, g.B)
            sub = fitting.FitResult(r.params[:rows], r.chi2_init[:rows], r.chi2[:rows], r.n_iter, r.n_accept[:rows],
                                    r.active[:rows], r.reason[:rows])
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nm = fitting.nm_fit_batch(g.lay, g.d_q, g.Iobs[:rows], g.start[:rows], g.free, eng=eng, return_device=True)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out[kind]["lm_same_systems"] = g.quality(sub, rows)
            out[kind]["nelder_mead"] = dict(g.quality(nm, rows), evaluations=int(nm.n_accept[0].item()), iterations=int(nm.n_iter),
                                            fits_per_s=rows / dt)
            a, b = out[kind]["lm_same_systems"], out[kind]["nelder_mead"]
            out[kind]["lm_at_least_nelder_mead"] = bool(a["success_frac"] >= b["success_frac"] and
                                                         a["median_chi2_over_truth"] <= b["median_chi2_over_truth"] * (1 + 1e-3))
    if with_nm:
        # not part of the timed fit and not what the reference does: the same LM loop after a coarse scan of the lattice
        # edge (fitting.scan_start), the failure mode of the crystal half, on the first `nm_rows` systems
        try:
            g = W.groups[1]
            rows = min(nm_rows, g.B)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            start = fitting.scan_start(g.lay, g.d_q, g.Iobs[:rows], g.start[:rows], g.lay.slot("x__a"), rel_window=0.12, n_points=25, eng=eng)
            rs = fitting.lm_fit_batch(g.lay, g.d_q, g.Iobs[:rows], start, g.free, eng=eng, return_device=True, max_iter=W.lm_iters)
            torch.cuda.synchronize()
            out["crystal"]["lm_after_lattice_scan"] = dict(g.quality(rs, rows), fits_per_s=rows / (time.perf_counter() - t0))
        except Exception as e:
            out["crystal"]["lm_after_lattice_scan"] = {"error": repr(e)}
    n = sum(g.B for g in W.groups)
    out["success_frac"] = sum(out[k]["lm"]["success_frac"] * out[k]["lm"]["systems"] for k in ("sphere", "crystal")) / n
    return out


def cpu_sample(workload, W, no_cpu, nm_evals=None):
    """cpu_baseline object of one workload: a bounded sample on the host cores (oracle/cpu_arm.py)."""
    if no_cpu:
        return None
    from oracle import cpu_arm
    cores = cpu_arm.host_cores()
    if workload in ("cfg2", "cfg3", "cfg4"):
        per_core = {"cfg2": 128, "cfg3": 2, "cfg4": 2}[workload]
        arm = cpu_arm.CpuArm(workload, cores)
        # the best of a few steps: a shared host shows 2x swings from one second to the next, and the baseline should be
        # the reference at its best on this box
        reps = 3 if workload == "cfg2" else 1
        n1, dt1 = min((arm.step(per_core, k * cores * per_core) for k in range(reps)), key=lambda r: r[1] / r[0])
        arm.close()
        val = n1 / dt1
        return {"value": val, "unit": UNITS[workload][1], "cores": cores, "kind": arm.kind, "per_core": val / cores,
                "sample": "%s%d parameter sets (%d per core) of the same workload through %s (reflection tables rebuilt per "
                          "pattern, as the reference does)%s" % ("best of %d steps of " % reps if reps > 1 else "", n1, per_core,
                                                              cpu_arm.describe(arm.kind),
                                                              "; one reference-equivalent evaluation = one pattern" if workload == "cfg4" else "")}
    # cfg5: sphere systems of this batch fitted on the host cores the reference's way, one per core
    g = W.groups[0]
    arm = cpu_arm.CpuArm("cfg5s", cores)
    try:
        rate_s, recs = arm.fits(g.start[:cores].cpu().numpy(), g.Iobs[:cores].cpu().numpy(), g.free)
    finally:
        arm.close()
    evals = float(np.median([r[1] for r in recs]))
    arm_x = cpu_arm.CpuArm("cfg4", cores)
    n, dtx = arm_x.step(1)
    arm_x.close()
    ev_x = float(nm_evals) if nm_evals else evals
    rate_x = (n / dtx) / ev_x
    mixed = 2.0 / (1.0 / rate_s + 1.0 / rate_x)
    return {"value": mixed, "unit": "fits/s", "cores": cores, "kind": arm.kind,
            "sphere_fits_per_s": rate_s, "crystal_fits_per_s_extrapolated": rate_x,
            "sphere_median_evaluations": evals, "sphere_median_chi2_ratio": float(np.median([r[3] / r[2] for r in recs])),
            "sample": "%d sphere systems of this batch fitted, one per core: scipy Nelder-Mead (xatol = fatol = 1e-4, at most 1500 "
                      "evaluations) on %s -- the reference's optimiser restated, lmfit itself is not installable here; crystal "
                      "half EXTRAPOLATED: %d cfg4 patterns timed (%.2f patterns/s) / %.0f evaluations per fit (the batched "
                      "Nelder-Mead count on the GPU for the same kind of system); 50/50 mix = harmonic mean"
                      % (len(recs), cpu_arm.describe(arm.kind), n, n / dtx, ev_x)}


def measure(eng, name, B, seed, steps, warmup, args, peaks, peaks_src, fp64, probe_ms, barrier=None, max_over_ranks=None,
            world=1, sampler=None, e_steps=None):
    """Everything the JSON line says about one workload: value, e2e, roofline, cpu_baseline (+ fit quality)."""
    import torch
    W = Work(eng, name, B, seed, args.lm_iters)
    for i in range(warmup):
        W.step(i)
    (barrier or torch.cuda.synchronize)()
    if sampler is not None:         # clocks and throttle reasons are sampled DURING the timed region only
        sampler.start()
    ms, launches = W.timed(steps, 0, barrier)
    clocks = sampler.result() if sampler is not None else None
    if max_over_ranks:
        ms = max_over_ranks(ms)
    value = world * W.units_per_step() * steps / (ms * 1e-3)
    fn, h2d, d2h, api, check = W.make_e2e()
    e_steps = e_steps or max(3, min(steps, 10))
    for _ in range(2):
        fn()
    (barrier or torch.cuda.synchronize)()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        fn()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if max_over_ranks:
        dt = max_over_ranks(dt)
    out = {"workload": WORKLOADS[name], "metric": W.metric, "unit": W.unit, "config": config_of(name, B), "value": value,
           "ms_per_step": ms / steps, "steps": steps, "gpu_launches": launches,
           "e2e": {"value": world * W.units_per_step() * e_steps / dt, "unit": W.unit, "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "steps": e_steps, "api": api},
           "checksum": check(), "q_evals_per_s": value * W.n_q, "clocks": clocks}
    return W, out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from xrsdkit_b200 import engine, batch as xb
    if world > 1:
        xb.bind_to_gpu_numa(local)      # host buffers of the e2e path next to this rank's GPU
    eng = engine.get_engine()
    dev = eng.device
    B = args.batch
    warmup = max(3, args.warmup)
    peaks, peaks_src = measured_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    fp64_tflops, probe_ms = eng.probe_fp64_peak()
    steps = args.steps if args.workload != "cfg5" else max(1, min(args.steps, 5))
    if args.workload == "cfg5":
        warmup = min(warmup, 3)
    W, head = measure(eng, args.workload, B, 1000 + rank, steps, warmup, args, peaks, peaks_src, fp64_tflops, probe_ms,
                      barrier, max_over_ranks, world, sampler=ClockSampler(local))
    clocks = head.pop("clocks")

    # ---- a fit whose results meet in one NCCL all_gather inside the timed region (N > 1)
    gathered = None
    if world > 1 and not args.no_secondary:
        gathered = gathered_fit(eng, args, rank, world, barrier, max_over_ranks)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    roofline = dominant_kernel(eng, W, fp64_tflops, probe_ms, peaks, peaks_src)
    quality = None
    nm_evals = None
    if args.workload == "cfg5":
        quality = fit_quality(eng, W, args.nm_rows, with_nm=not args.no_nm)
        nm_evals = (quality.get("crystal", {}).get("nelder_mead") or {}).get("evaluations")
    cpu = cpu_sample(args.workload, W, args.no_cpu or world > 1, nm_evals)

    # ---- the other BASELINE configs at full batch (N = 1, default workload only)
    secondary = None
    if args.workload == "cfg2" and world == 1 and not args.no_secondary:
        secondary = {}
        outs0 = W.outs[0]
        for name, b2, st, wu in (("cfg3", args.batch, 5, 3), ("cfg4", args.batch, 3, 3), ("cfg5", args.fit_batch, 2, 2)):
            try:
                w2, o2 = measure(eng, name, b2, 2000, st, wu, args, peaks, peaks_src, fp64_tflops, probe_ms,
                                 sampler=ClockSampler(local))
                o2["roofline"] = dominant_kernel(eng, w2, fp64_tflops, probe_ms, peaks, peaks_src)
                ev = None
                if name == "cfg5":
                    o2["fit_quality"] = fit_quality(eng, w2, args.nm_rows, with_nm=not args.no_nm)
                    ev = (o2["fit_quality"].get("crystal", {}).get("nelder_mead") or {}).get("evaluations")
                try:
                    o2["cpu_baseline"] = cpu_sample(name, w2, args.no_cpu, ev)
                except Exception as e:
                    o2["cpu_baseline"] = {"error": repr(e)}
                if o2.get("cpu_baseline") and o2["cpu_baseline"].get("value"):
                    o2["speedup_vs_cpu"] = {"device": o2["value"] / o2["cpu_baseline"]["value"],
                                            "e2e": o2["e2e"]["value"] / o2["cpu_baseline"]["value"]}
                secondary[name] = o2
                del w2
                torch.cuda.empty_cache()
            except Exception as e:      # a secondary measurement must not take the headline down
                secondary[name] = {"error": repr(e)}
        secondary["cfg2_fp32"] = bench_fp32(eng, W)
        secondary["features"] = bench_features(eng, W, outs0, args)
        secondary["cfg1"] = bench_cfg1(args)
        secondary["cfg1_batched"] = bench_cfg1_batched(eng, peaks, peaks_src, fp64_tflops)

    line = {
        "metric": head["metric"], "value": head["value"], "unit": head["unit"], "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": head["config"],
        "e2e": head["e2e"],
        "gpu_launches": head["gpu_launches"],
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "details": {"q_evals_per_s": head["q_evals_per_s"], "checksum": head["checksum"],
                    "evaluations_note": ("reference-equivalent: a forward-difference Jacobian of 8 parameters costs the "
                                         "reference 9 pattern evaluations per system; %d are executed here, the I0 "
                                         "derivatives are analytic" % W.n_var) if args.workload == "cfg4" else None},
        "fit_quality": quality,
        "secondary": secondary,
    }
    if gathered is not None:
        line["secondary"] = dict(line["secondary"] or {}, cfg5_gathered=gathered)
    write_fp64_peak(fp64_tflops, probe_ms, clocks)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def gathered_fit(eng, args, rank, world, barrier, max_over_ranks):
    """cfg5 sharded over the ranks with the final NCCL gather of (fitted parameters, chi2) INSIDE the timed region
    (BASELINE.json north_star: "using only a final NCCL gather of results over NVLink")."""
    import torch
    from xrsdkit_b200 import batch as xb
    per_rank = args.gather_batch
    W = Work(eng, "cfg5", per_rank, 3000 + rank, args.lm_iters)
    W.one_pass(0)               # warm-up fit (allocations, table stores)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    W.one_pass(0)
    full = []
    for g, r in zip(W.groups, W.res):
        rows = torch.cat([r.params, r.chi2[:, None]], dim=1)
        full.append((rows, xb.gather_rows(rows, world * g.B)))
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    ok = all(bool(torch.equal(all_rows[rank * rows.shape[0]:(rank + 1) * rows.shape[0]], rows)) for rows, all_rows in full)
    ok_all = max_over_ranks(0.0 if ok else 1.0) == 0.0
    q = [g.quality(r) for g, r in zip(W.groups, W.res)]
    return {"workload": WORKLOADS["cfg5"], "unit": "fits/s", "value": world * per_rank / (ms * 1e-3), "ms": ms,
            "fits_per_rank": per_rank, "gathered_rows": [int(a.shape[0]) for _, a in full],
            "gather_bytes_per_rank": int(sum(r.numel() * 8 for r, _ in full)),
            "own_shard_equals_gathered_rows_on_every_rank": ok_all,
            "success_frac_rank0": sum(x["success_frac"] * x["systems"] for x in q) / sum(x["systems"] for x in q),
            "collective": "torch.distributed.all_gather_into_tensor (NCCL), inside the timed region"}


def bench_features(eng, W, pats0, args):
    import torch
    try:        # featuriser over this workload's own patterns, with 2 % noise
        g = torch.Generator(eng.device).manual_seed(7)
        pats = pats0 * (1 + 0.02 * torch.randn(pats0.shape, dtype=torch.float64, device=eng.device, generator=g))
        f_out = eng.features(W.d_q, pats)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.features(W.d_q, pats, out=f_out)
        e1.record()
        torch.cuda.synchronize()
        msf = e0.elapsed_time(e1) / 3
        out = {"workload": "profile_pattern (21 features) of %d noisy cfg2 patterns, n_q=%d" % (W.B, W.n_q),
               "unit": "patterns/s", "value": W.B / (msf * 1e-3), "ms_per_step": msf,
               "read_GBps": W.B * W.n_q * 8 / (msf * 1e-3) / 1e9}
        if not args.no_cpu:
            from oracle import xrsd_oracle as xo
            hp = pats[:2].cpu().numpy()
            t0 = time.perf_counter()
            ref_f = np.stack([xo.profile_pattern(W.q, row) for row in hp])
            out["cpu_patterns_per_s_1core"] = 2 / (time.perf_counter() - t0)
            out["max_rel_diff_vs_oracle"] = float(np.max(np.abs(f_out[:2].cpu().numpy() - ref_f)[:, :13] / np.abs(ref_f[:, :13])))
        return out
    except Exception as e:
        return {"error": repr(e)}


def bench_fp32(eng, W):
    """The optional single-precision path on the headline workload: same batch, FP32 size-distribution loops,
    kernel-only rate and the largest relative difference to the FP64 result (the gate is 1e-5)."""
    import torch
    try:
        ref = eng.intensity(W.lay, W.d_q, W.d_params[0]).clone()
        out = torch.empty_like(ref)
        eng.intensity(W.lay, W.d_q, W.d_params[0], out=out, precision="fp32")
        diff = float(((out - ref).abs() / ref.abs()).max().item())
        ms = time_kernel(lambda i: eng.intensity(W.lay, W.d_q, W.d_params[i % 4], out=W.outs[i % 2], precision="fp32"))
        ms64 = time_kernel(lambda i: eng.intensity(W.lay, W.d_q, W.d_params[i % 4], out=W.outs[i % 2]))
        return {"workload": WORKLOADS["cfg2"] + " -- optional FP32 path (xrsd_intensity_batch_f32)", "unit": "patterns/s",
                "value": W.B / (ms * 1e-3), "kernel_ms": ms, "fp64_kernel_ms": ms64, "speedup_over_fp64": ms64 / ms,
                "max_rel_diff_vs_fp64": diff, "gate": 1e-5, "dtype": "f32 loops, f64 per-q work and output"}
    except Exception as e:
        return {"error": repr(e)}


def bench_cfg1(args):
    try:        # BASELINE configs[0]: ONE pattern through System.compute_intensity (numpy in, numpy out)
        from xrsdkit_b200.system import System
        s1 = System(p=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}, "I0": {"value": 1.0}}))
        q1 = np.linspace(0.001, 0.5, 2000)
        for _ in range(200):
            I1 = s1.compute_intensity(q1)
        t0 = time.perf_counter()
        for _ in range(2000):
            I1 = s1.compute_intensity(q1)
        us1 = (time.perf_counter() - t0) / 2000 * 1e6
        out = {"workload": "cfg1: single dilute spherical population (r=20), 2000-pt q, ONE pattern per call "
                           "through System.compute_intensity with host arrays (latency-bound)",
               "unit": "us per call", "value": us1, "higher_is_better": False}
        if not args.no_cpu:
            from oracle import cpu_arm
            be = cpu_arm.Backend("cfg1")
            row = be.layout.values
            for _ in range(50):
                Ir = be.intensity(row)
            t0 = time.perf_counter()
            for _ in range(500):
                Ir = be.intensity(row)
            out["cpu_us_per_call_1core"] = (time.perf_counter() - t0) / 500 * 1e6
            out["cpu_kind"] = be.kind
            out["max_rel_diff_vs_cpu"] = float(np.max(np.abs(I1 - Ir) / np.abs(Ir)))
        return out
    except Exception as e:
        return {"error": repr(e)}


def bench_cfg1_batched(eng, peaks, peaks_src, fp64_tflops, B=100000):
    """SURVEY 8d's batched variant of cfg1: B = 10^5 dilute single-size sphere systems, r ~ U(5, 60), 2000 q -- about
    60 flop (one sincos) per 8-byte I(q) value written, i.e. the one workload of the path near the HBM / FP64 ridge
    instead of far on the compute side; reported against both ceilings."""
    import torch
    try:
        from xrsdkit_b200.system import System
        from xrsdkit_b200 import batch as xb
        s1 = System(p=dict(structure="diffuse", form="spherical", parameters={"r": {"value": 20.0}, "I0": {"value": 1.0}}))
        q1 = np.linspace(0.001, 0.5, 2000)
        lay = xb.layout_of(s1, q1)
        rs = np.random.RandomState(1001)
        mats = []
        for _ in range(2):
            P = np.tile(lay.values, (B, 1))
            P[:, lay.slot("p__r")] = rs.uniform(5.0, 60.0, B)
            mats.append(eng.to_device(P))
        d_q = eng.to_device(q1)
        outs = [torch.empty((B, len(q1)), dtype=torch.float64, device=eng.device) for _ in range(2)]
        ms = time_kernel(lambda i: eng.intensity(lay, d_q, mats[i % 2], out=outs[i % 2]), reps=5)
        nbytes = 8.0 * B * len(q1)
        # parity of a few rows against the closed form the reference evaluates (form_factors.py:20-28), in numpy
        rows = [0, B // 2, B - 1]
        got = eng.intensity(lay, d_q, mats[0][rows].contiguous()).cpu().numpy()
        r = mats[0][rows, lay.slot("p__r")].cpu().numpy()[:, None]
        x = q1[None, :] * r
        ff = 3.0 * (np.sin(x) - x * np.cos(x)) / x ** 3
        ref = float(lay.values[lay.slot("noise__I0")]) + 1.0 * ff ** 2
        gbs = nbytes / (ms * 1e-3) / 1e9
        return {"workload": "cfg1 batched: %d dilute sphere systems (r ~ U(5, 60)), 2000-pt q, one launch; every pass writes "
                            "%.1f GB of patterns (L2 is 126 MB)" % (B, nbytes / 1e9),
                "unit": "patterns/s", "value": B / (ms * 1e-3), "kernel_ms": ms,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                             "frac": gbs / peaks.get("hbm_gbs", 6650.0), "peak_source": peaks_src,
                             "algorithmic_bytes_per_launch": nbytes,
                             "note": "write-only stream; MEASURED_PEAKS' figure is a copy (read + write)",
                             "fp64_frac_survey_convention": 60.0 * B * len(q1) / (ms * 1e-3) / 1e12 / fp64_tflops},
                "max_rel_diff_vs_closed_form": float(np.max(np.abs(got - ref) / np.abs(ref)))}
    except Exception as e:
        return {"error": repr(e)}


def write_fp64_peak(tflops, probe_ms, clocks):
    """The roofline denominator of this run, with the clock samples taken beside it (profiles/fp64_peak.json is the
    committed copy of one such run; DESIGN.md section 4 has the SASS-level justification of the probe)."""
    rec = {"probe_tflops": tflops, "probe_ms": probe_ms, "nominal_tflops": FP64_NOMINAL_TFLOPS,
           "probe_over_nominal": tflops / FP64_NOMINAL_TFLOPS, "clocks": clocks,
           "probe": "fp64_probe_kernel: 1184 CTAs x 256 threads x 8 independent chains x 32768 DFMA; the SASS loop is 8 DFMA per "
                    "iteration unrolled (cuobjdump -sass: DFMA only, plus the loop counter), i.e. 2 flop per issued FP64 instruction"}
    for d in ("gpurun_out", "profiles"):
        p = os.path.join(ROOT, d)
        if os.path.isdir(p):
            try:
                with open(os.path.join(p, "fp64_peak.json"), "w") as f:
                    json.dump(rec, f, indent=1)
                break
            except Exception:
                pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=10000, help="parameter sets (systems) per GPU per pass")
    ap.add_argument("--fit-batch", type=int, default=16384, help="systems of the secondary cfg5 measurement (half sphere, half crystal)")
    ap.add_argument("--gather-batch", type=int, default=125000,
                    help="fits per rank of the gathered cfg5 measurement (N > 1); the default makes 8 ranks BASELINE config 5: 10^6 fits")
    ap.add_argument("--nm-rows", type=int, default=1024, help="systems per kind also fitted by batched Nelder-Mead (quality comparison)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline samples")
    ap.add_argument("--no-nm", action="store_true", help="skip the Nelder-Mead comparison of the fit workloads")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg3/cfg4/cfg5 measurements")
    ap.add_argument("--lm-iters", type=int, default=LM_ITERS, help="LM iteration cap per fit (cfg5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
