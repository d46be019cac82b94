#!/usr/bin/env python
"""bench.py -- I(q) pattern evaluations per second on B200, beside the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5]
    python bench.py --impl reference ...        # the CPU arm (oracle port of the reference's numpy path)

A *step* is one pass of the hot path over one batch of synthetic parameter sets of the
workload (default: BASELINE.json configs[1] = "cfg2": polydisperse spheres, 50-pt r_normal
quadrature, hard-sphere S(q), 4096-pt q grid, 10k parameter sets per GPU).  Under torchrun
every rank runs the same per-GPU batch (weak scaling); rank 0 prints ONE JSON line.

value       patterns/s with inputs resident in HBM (CUDA events around K steps, max over ranks)
e2e         patterns/s through the public host-buffer API: pinned host params in, I(q) out to
            pinned host memory, copies inside the timed region
roofline    dominant kernel against the FP64 (DFMA) pipe peak measured live by a probe kernel,
            with achieved HBM GB/s against MEASURED_PEAKS.json beside it
cpu_baseline the oracle (numpy restatement of the reference) on the host cores, bounded sample
"""
import argparse
import json
import multiprocessing as mp
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

# Algorithmic FP64 work per unit, as stated in DESIGN.md section 4 (FMA = 2, mul/add = 1, div = 10,
# sincos pair = 40, exp = 30, Re w = the series actually needed):
#   cfg2: per (set, q): n_r * 13 (rotation 4 ops = 6 flop, x 1, t 2, t^2 1, acc 2 ... = 13) + 250
FLOPS_PER_PAIR_RNORMAL = 13.0
FLOPS_PER_Q_OVERHEAD_CFG2 = 250.0
SURVEY_FLOPS_PER_SETQ_CFG2 = 2630.0          # SURVEY.md 8d convention (sincos per (q,r) pair)


# --------------------------------------------------------------------------------------
# workloads (SURVEY.md 8d synthetic inputs; seeds 1000 + cfg#)
# --------------------------------------------------------------------------------------
def make_workload(name, batch):
    from helpers import cfg2_system, cfg2_params, cfg3_system, cfg3_params, cfg4_system, cfg4_params
    from xrsdkit_b200.system import System
    from xrsdkit_b200 import batch as xb
    if name == "cfg2":
        s = cfg2_system(System)
        q = np.linspace(0.01, 0.5, 4096)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg2_params(lay, batch)
    if name == "cfg3":
        s = cfg3_system(System)
        q = np.linspace(0.01, 1.0, 8192)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg3_params(lay, batch)
    if name in ("cfg4", "cfg5"):
        s = cfg4_system(System)
        q = np.linspace(0.01, 1.0, 8192)
        lay = xb.layout_of(s, q)
        return s, q, lay, cfg4_params(lay, batch)
    raise SystemExit("unknown workload " + name)


# --------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# --------------------------------------------------------------------------------------
_W = {}


def _cpu_init(workload):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    from oracle import xrsd_oracle as orc
    from helpers import params_to_sysdict
    from xrsdkit_b200 import definitions as d
    s, q, lay, P = make_workload(workload, 4096 if workload == "cfg2" else 256)
    _W.update(orc=orc, s=s, q=q, lay=lay, P=P, to_dict=params_to_sysdict,
              tables=dict(atomic_table=d.atomic_params, sg_point_groups=d.sg_point_groups,
                          lattice_space_groups=d.lattice_space_groups))


def _cpu_eval(args):
    lo, hi = args
    w = _W
    acc = 0.0
    for b in range(lo, hi):
        row = w["P"][b % len(w["P"])]
        I = w["orc"].system_intensity(w["q"], w["to_dict"](w["s"], w["lay"], row), **w["tables"])
        acc += float(I[0])
    return hi - lo, acc


class CpuArm(object):
    def __init__(self, workload, cores):
        self.cores = cores
        self.pool = mp.get_context("spawn").Pool(cores, initializer=_cpu_init, initargs=(workload,))
        self.pool.map(_cpu_eval, [(0, 1)] * cores)          # import + first-touch warm-up

    def step(self, sets_per_core, offset=0):
        t0 = time.perf_counter()
        jobs = [(offset + c * sets_per_core, offset + (c + 1) * sets_per_core) for c in range(self.cores)]
        done = sum(n for n, _ in self.pool.map(_cpu_eval, jobs))
        return done, time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def host_cores():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    per_core = {"cfg2": 8, "cfg3": 1, "cfg4": 1, "cfg5": 1}[args.workload]
    arm = CpuArm(args.workload, cores)
    for w in range(args.warmup):
        arm.step(per_core, w * cores * per_core)
    t_tot, n_tot = 0.0, 0
    for k in range(args.steps):
        n, dt = arm.step(per_core, (args.warmup + k) * cores * per_core)
        t_tot += dt
        n_tot += n
    arm.close()
    val = n_tot / t_tot
    nq = 4096 if args.workload == "cfg2" else 8192
    line = {
        "impl": "reference", "metric": "I(q) pattern evals/sec", "value": val, "unit": "patterns/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "n_q": nq, "sets_per_step": cores * per_core,
                   "q_evals_per_s": val * nq},
        "cpu_baseline": {"value": val, "unit": "patterns/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d parameter sets (%d per core) of the same synthetic workload, "
                                   "numpy oracle port of the reference (oracle/xrsd_oracle.py), one process per core"
                                   % (args.steps, cores * per_core, per_core)},
        "e2e": {"value": val, "unit": "patterns/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


def workload_name(w):
    return {
        "cfg2": "cfg2: polydisperse spheres (r_normal, 50-pt quadrature) + hard_spheres S(q), 4096-pt q, 10k parameter sets per GPU",
        "cfg3": "cfg3: crystalline FCC Fm-3m, spherical form, voigt, q_max=1.0, 8192-pt q, table build + profile sum",
        "cfg4": "cfg4: FCC crystal + guinier_porod + flat noise, 8192-pt q, residual + FD Jacobian (8 free parameters)",
        "cfg5": "cfg5: batched Levenberg-Marquardt fit of cfg4-like crystal systems",
    }[w]


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
    eng = engine.get_engine()
    dev = eng.device
    B = args.batch
    s, q, lay, P = make_workload(args.workload, B)
    n_q = len(q)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    d_q = eng.to_device(q)
    # rotate over 4 distinct parameter matrices and 2 output buffers; the 328 MB output of one step
    # already exceeds the 126 MB L2, so no step finds its data cached
    rs = np.random.RandomState(1000 + rank)
    d_params = [eng.to_device(P[rs.permutation(B)]) for _ in range(4)]
    outs = [torch.empty((B, n_q), dtype=torch.float64, device=dev) for _ in range(2)]
    from xrsdkit_b200 import lowering
    max_cells = lowering.max_cells(lay, P) if lay.n_xtal else None
    tables = [eng.build_tables(lay, p, B, params_host=P) if lay.n_xtal else None for p in d_params]

    def step(i):
        if lay.n_xtal and args.workload == "cfg3":
            t = eng.build_tables(lay, d_params[i % 4], B, max_cells=max_cells, check=False, reuse=tables[i % 4])   # table build is part of the step
        else:
            t = tables[i % 4]
        eng.intensity(lay, d_q, d_params[i % 4], tables=t, out=outs[i % 2])

    for i in range(max(3, args.warmup)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.result()
    launches = eng.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)

    # ---- e2e through the host-buffer API (pinned host memory, copies timed)
    p_host = torch.as_tensor(P).pin_memory()
    q_host = torch.as_tensor(q).pin_memory()
    I_host = torch.empty((B, n_q), dtype=torch.float64).pin_memory()
    for _ in range(2):
        xb.compute_intensity_pinned(lay, q_host, p_host, I_host, eng=eng)
    barrier()
    e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        xb.compute_intensity_pinned(lay, q_host, p_host, I_host, eng=eng)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    te = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * B * e_steps / float(te.item())
    checksum = float(I_host[:: max(1, B // 7)].sum().item())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (intensity_kernel), timed alone with CUDA events
    peaks, peaks_src = measured_peaks()
    fp64_tflops, probe_ms = eng.probe_fp64_peak()
    for i in range(3):
        eng.intensity(lay, d_q, d_params[i % 4], tables=tables[i % 4], out=outs[i % 2])
    torch.cuda.synchronize()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    k0.record()
    for i in range(reps):
        eng.intensity(lay, d_q, d_params[i % 4], tables=tables[i % 4], out=outs[i % 2])
    k1.record()
    torch.cuda.synchronize()
    k_ms = k0.elapsed_time(k1) / reps
    if args.workload == "cfg2":
        n_r = 50.5     # 2*2.5/0.1 = 50 exactly: numpy.arange yields 50 or 51 radii depending on rounding
        flops_unit = FLOPS_PER_PAIR_RNORMAL * n_r + FLOPS_PER_Q_OVERHEAD_CFG2
        survey_unit = SURVEY_FLOPS_PER_SETQ_CFG2
    else:
        ns = float(np.mean([int(t_.n_shell.float().mean().item()) for t_ in tables if t_ is not None]))
        flops_unit = 60.0 * ns + 150.0           # far-field Voigt ~60 flop per (q, shell)
        survey_unit = 205.0 * ns + 150.0
    flops = flops_unit * B * n_q
    achieved = flops / (k_ms * 1e-3) / 1e12
    hbm_bytes = 8.0 * B * n_q                     # the pattern is written once; q and params stay in L2
    roofline = {
        "bound": "fp64", "achieved": achieved, "peak": fp64_tflops, "unit": "TFLOP/s", "frac": achieved / fp64_tflops,
        "traffic": None,
        "kernel": "xrsd::intensity_kernel<2,false>", "kernel_ms": k_ms,
        "flops_per_unit": flops_unit, "units_per_launch": B * n_q,
        "peak_source": "xrsd_probe_fp64_peak: DFMA-bound kernel timed in this run (%.3f ms); MEASURED_PEAKS.json has no FP64 figure" % probe_ms,
        "achieved_survey_convention": survey_unit * B * n_q / (k_ms * 1e-3) / 1e12,
        "hbm": {"achieved": hbm_bytes / (k_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": hbm_bytes / (k_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs", 6650.0), "peak_source": peaks_src},
    }

    # ---- CPU baseline: bounded sample on the host cores
    cores = host_cores()
    per_core = {"cfg2": 24, "cfg3": 2, "cfg4": 2, "cfg5": 2}[args.workload]
    cpu = None
    if not args.no_cpu:
        arm = CpuArm(args.workload, cores)
        n1, dt1 = arm.step(per_core)
        n2, dt2 = arm.step(per_core, cores * per_core)
        arm.close()
        cpu_val = (n1 + n2) / (dt1 + dt2)
        cpu = {"value": cpu_val, "unit": "patterns/s", "cores": cores, "kind": "port",
               "sample": "2 x %d parameter sets (%d per core) of the same workload, numpy oracle port of the "
                         "reference's path (reflection tables rebuilt per pattern, as the reference does)" % (cores * per_core, per_core),
               "per_core": cpu_val / cores}

    line = {
        "metric": "I(q) pattern evals/sec", "value": value, "unit": "patterns/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload), "n_q": n_q, "sets_per_gpu_per_step": B,
                   "q_evals_per_s": value * n_q,
                   "l2": "each step writes %.0f MB (> 126 MB L2) and rotates over 4 parameter matrices" % (8e-6 * B * n_q),
                   "checksum": checksum},
        "e2e": {"value": e2e, "unit": "patterns/s", "h2d_bytes_per_step": int(P.nbytes + q.nbytes),
                "d2h_bytes_per_step": int(8 * B * n_q), "steps": e_steps,
                "api": "xrsdkit_b200.batch.compute_intensity_pinned (pinned host buffers, chunked copy/compute overlap)"},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--batch", type=int, default=10000, help="parameter sets per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline sample")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
