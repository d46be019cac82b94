#!/usr/bin/env python
"""Fit-quality probe: batched LM against batched Nelder-Mead (the reference's optimiser) on the SAME synthetic
systems, SURVEY.md 8d cfg5 settings (I_obs = I(truth) (1 + 0.02 randn), start = truth (1 + jitter U(-1,1))).

    python profiles/fit_quality_probe.py [--batch 2048] [--jitter 0.1] [--out gpurun_out/fit_quality.json]

Success: chi2_final <= chi2(truth) * (1 + eps).  Prints one JSON object per (layout, method, setting).
Needs a GPU; used to choose the LM loop's defaults (iteration cap, relative trust region) -- not a test.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2048)
    ap.add_argument("--jitter", type=float, nargs="+", default=[0.03, 0.10])
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-nm", action="store_true")
    args = ap.parse_args()
    import torch
    from helpers import cfg2_system, cfg2_params, cfg4_system, cfg4_params
    from xrsdkit_b200 import engine, fitting, batch as xb
    from xrsdkit_b200.system import System
    eng = engine.get_engine()
    dev = eng.device
    rows = []

    def emit(**kw):
        rows.append(kw)
        print(json.dumps(kw), flush=True)

    def quality(res, chi2_truth):
        r = (res.chi2 / chi2_truth).double()
        r = torch.where(torch.isfinite(r), r, torch.full_like(r, float("inf")))
        qs = torch.quantile(r[torch.isfinite(r)], torch.tensor([0.1, 0.5, 0.9, 0.99], dtype=torch.float64, device=dev)).tolist() \
            if torch.isfinite(r).any() else [float("nan")] * 4
        return dict(success_1pct=float((r <= 1.01).double().mean().item()), success_10pct=float((r <= 1.10).double().mean().item()),
                    success_2x=float((r <= 2.0).double().mean().item()),
                    q10=qs[0], q50=qs[1], q90=qs[2], q99=qs[3],
                    median_chi2_ratio_init=float((res.chi2 / res.chi2_init).median().item()))

    for lname in ("sphere", "crystal"):
        if lname == "sphere":
            s, q = cfg2_system(System), np.linspace(0.01, 0.5, 4096)
            lay = xb.layout_of(s, q)
            P = cfg2_params(lay, args.batch)
            names = ("noise__I0", "p__I0", "p__r", "p__sigma", "p__r_hard", "p__v_fraction")
        else:
            s, q = cfg4_system(System), np.linspace(0.01, 1.0, 8192)
            lay = xb.layout_of(s, q)
            P = cfg4_params(lay, args.batch)
            names = ("noise__I0", "x__I0", "x__a", "x__hwhm_g", "x__hwhm_l", "x__r", "gp__I0", "gp__rg")
        free = [lay.slot(k) for k in names]
        d_q = eng.to_device(q)
        truth = eng.to_device(P)
        g = torch.Generator(dev).manual_seed(4242)
        Iobs = eng.intensity(lay, d_q, truth).clone()
        Iobs *= 1 + 0.02 * torch.randn(Iobs.shape, dtype=torch.float64, device=dev, generator=g)
        obj = eng.objective(True, True, (0.0, float("inf")))
        chi2_truth = eng.residual(lay, obj, d_q, truth, Iobs).clone()
        mask = torch.zeros(lay.n_params, dtype=torch.bool, device=dev)
        mask[free] = True
        for jit in args.jitter:
            u = 2 * torch.rand(truth.shape, dtype=torch.float64, device=dev, generator=g) - 1
            start = torch.where(mask[None, :], truth * (1 + jit * u), truth).contiguous()
            # keep v_fraction inside its bounds (0.01..0.7405) as lmfit would clip the start value
            lo, hi = lay.bounds(list(range(lay.n_params)))
            start = torch.minimum(torch.maximum(start, eng.to_device(np.where(np.isfinite(lo), lo, -1e300))[None, :]),
                                  eng.to_device(np.where(np.isfinite(hi), hi, 1e300))[None, :]).contiguous()
            chi2_start = eng.residual(lay, obj, d_q, start, Iobs).clone()
            emit(layout=lname, jitter=jit, method="start", batch=args.batch,
                 median_chi2_start_over_truth=float((chi2_start / chi2_truth).median().item()),
                 median_chi2_truth=float(chi2_truth.median().item()))
            for iters in (10, 20, 30, 60, 120):
                for cap in (0.0, 0.5, 0.25):
                    for lam0 in (1e-2, 1.0):
                        if (cap, lam0) not in ((0.0, 1e-2), (0.5, 1e-2), (0.25, 1e-2), (0.5, 1.0)):
                            continue
                        if iters in (10, 120) and (cap, lam0) != (0.5, 1e-2):
                            continue
                        torch.cuda.synchronize()
                        t0 = time.perf_counter()
                        res = fitting.lm_fit_batch(lay, d_q, Iobs, start, free, max_iter=iters, eng=eng, return_device=True,
                                                   max_rel_step=cap, lambda0=lam0)
                        torch.cuda.synchronize()
                        dt = time.perf_counter() - t0
                        emit(layout=lname, jitter=jit, method="lm", iters=iters, max_rel_step=cap, lambda0=lam0,
                             seconds=dt, fits_per_s=args.batch / dt, still_active=float(res.active.double().mean().item()),
                             mean_accepts=float(res.n_accept.double().mean().item()), **quality(res, chi2_truth))
            if not args.no_nm:
                for mi in (None,):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    res = fitting.nm_fit_batch(lay, d_q, Iobs, start, free, eng=eng, return_device=True, max_iter=mi)
                    torch.cuda.synchronize()
                    dt = time.perf_counter() - t0
                    emit(layout=lname, jitter=jit, method="nelder-mead", iters=res.n_iter, evaluations=int(res.n_accept[0].item()),
                         seconds=dt, fits_per_s=args.batch / dt, not_converged=float(res.active.double().mean().item()),
                         **quality(res, chi2_truth))
    if args.out:
        with open(args.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
