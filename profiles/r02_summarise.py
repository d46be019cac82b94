#!/usr/bin/env python
"""Summarise the .ncu-rep captures of profiles/r02_capture.sh (brought back in gpurun_out/) into the text files and
the counts file kept under profiles/:

    python profiles/r02_summarise.py            # -> profiles/r02_ncu_<cfg>.txt, profiles/r02_ncu_counts.json

Per kernel launch: duration, DRAM bytes, registers, occupancy limits, warps active, FP64 pipe utilisation, issue
utilisation, executed instructions, local loads / stores, and the executed FP64 thread instructions
(DFMA / DADD / DMUL, predicated-on) from which the executed flops follow (2 per DFMA, 1 per DADD / DMUL) -- the
figure bench.py reports beside its hand-counted `flops_per_unit` (roofline.ncu)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.per_cycle_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg.per_second",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads",
    "sass__inst_executed_local_stores", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__inst_executed_pipe_fp64.sum",
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "Tbyte": 1e12}


def raw_page(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[head], rows[head + 1]
    return names, units, rows[head + 2:]


def summarise(tag, title):
    rep = os.path.join(OUT, "r02_%s.ncu-rep" % tag)
    if not os.path.exists(rep):
        return None
    names, units, launches = raw_page(rep)
    col = {n: i for i, n in enumerate(names)}
    lines = ["# Round 2: %s" % title, "# source: gpurun_out/r02_%s.ncu-rep (ncu --set full --clock-control none), profiles/r02_capture.sh" % tag]
    recs = []
    kn = [r[col["Kernel Name"]][:60] for r in launches]
    lines.append("%-72s %s" % ("Kernel Name", " | ".join(kn)))
    for m in KEEP:
        if m not in col:
            continue
        lines.append("%-72s %-16s %s" % (m, units[col[m]], " | ".join(r[col[m]] for r in launches)))
    for r in launches:
        def val(m):
            try:
                return float(r[col[m]].replace(",", ""))
            except Exception:
                return None
        stalls = sorted(((float(r[i].replace(",", "") or 0), n[len(STALLS):]) for n, i in col.items()
                         if n.startswith(STALLS) and not n.endswith("_not_issued") and r[i]), reverse=True)
        tot = sum(s for s, _ in stalls) or 1.0
        lines.append("# stall samples %s: %s" % (r[col["Kernel Name"]][:48], ", ".join("%s %.0f%%" % (n, 100 * s / tot) for s, n in stalls[:7])))
        dfma, dadd, dmul = (val("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % k) for k in ("dfma", "dadd", "dmul"))
        rec = {"kernel": r[col["Kernel Name"]], "ms": val("gpu__time_duration.sum"),
               "dram_bytes": (val("dram__bytes_read.sum") or 0) * UNIT.get(units[col["dram__bytes_read.sum"]], 1.0) +
                             (val("dram__bytes_write.sum") or 0) * UNIT.get(units[col["dram__bytes_write.sum"]], 1.0),
               "fp64_pipe_pct": val("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
               "grid": val("launch__grid_size")}
        if dfma is not None:
            rec.update(dfma=dfma, dadd=dadd, dmul=dmul, fp64_flops=2 * dfma + (dadd or 0) + (dmul or 0))
            lines.append("# executed FP64 thread instructions %s: DFMA %.4g, DADD %.4g, DMUL %.4g -> %.4g flop per launch" %
                         (r[col["Kernel Name"]][:40], dfma, dadd or 0, dmul or 0, rec["fp64_flops"]))
        recs.append(rec)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_%s.txt" % tag), "w") as f:
        f.write("\n".join(lines) + "\n")
    return recs


def main():
    counts = {}
    for tag, title, pick in (("cfg2", "cfg2, 10k sets x 4096 q: intensity_kernel<2,false,false>", "intensity_kernel"),
                             ("cfg3", "cfg3, 10k FCC patterns x 8192 q: table_key_kernel + reflection_table_kernel + intensity_kernel<2,false,true>", "intensity_kernel"),
                             ("cfg4", "cfg4, 10k systems x 8192 q: jacobian_variants_kernel + jacobian_reduce_kernel", "jacobian_variants")):
        recs = summarise(tag, title)
        if not recs:
            continue
        top = [r for r in recs if pick in r["kernel"]]
        if top:
            r = top[0]
            # CTAs per system of the captured launch: one whole-pattern CTA per pattern (cfg2, cfg3), one per executed
            # variant (cfg4: base + 5 non-linear parameters)
            per_sys = {"cfg2": 1, "cfg3": 1, "cfg4": 6}[tag]
            counts[tag] = {"dram_bytes": r["dram_bytes"], "fp64_flops": r.get("fp64_flops"), "fp64_pipe_pct": r["fp64_pipe_pct"],
                           "kernel_ms_under_ncu": r["ms"], "source": "profiles/r02_ncu_%s.txt" % tag,
                           "systems_per_launch": int(round(r["grid"] / per_sys)),
                           "all": [{k: v for k, v in x.items()} for x in recs]}
    if "cfg4" in counts:
        counts["cfg5"] = dict(counts["cfg4"], source=counts["cfg4"]["source"] + " (same kernel; the cfg4 launch shape)")
        counts["cfg5"].pop("dram_bytes", None)
        counts["cfg5"].pop("fp64_flops", None)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_counts.json"), "w") as f:
        json.dump(counts, f, indent=1)
    print(json.dumps({k: {kk: vv for kk, vv in v.items() if kk != "all"} for k, v in counts.items()}, indent=1))


if __name__ == "__main__":
    sys.exit(main())
