#!/bin/bash
# Round-2 final-tree evidence (15-node far-field interpolant), run on the GPU box from the repo root:
#   gpurun --timeout 900 -- 'bash profiles/r02d_capture.sh [tag]'      (tag: prefix of the bench / test files, default r02d)
# GPU tests, the default bench line, the reference arm, then the cfg3 / cfg4 captures again (same commands as
# profiles/r02_capture.sh: plain run first, must exit 0, only then the launch list and the ncu --set full pass).
# Outputs land in gpurun_out/ and are summarised into profiles/ by profiles/r02_summarise.py.
set -u
O=gpurun_out; TAG=${1:-r02d}
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.txt 2>&1; echo "pytest rc=$?"
python bench.py > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_reference.json 2> $O/${TAG}_bench_reference.err; echo "reference rc=$?"
FP64="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_fp64.sum"
NCU="ncu --set full --metrics $FP64 --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"
C3="python bench.py --workload cfg3 --steps 2 --warmup 3 --no-cpu --no-secondary"
$C3 > $O/r02_plain_cfg3.json 2> $O/r02_plain_cfg3.err && {
  $LIST -c 400 --log-file $O/r02_launches_cfg3.csv $C3 > $O/r02_ncu_list_cfg3.log 2>&1
  $NCU -k "regex:intensity_kernel|reflection_table|table_key" -s 60 -c 3 -f -o $O/r02_cfg3 $C3 > $O/r02_ncu_cfg3.log 2>&1
}
C4="python bench.py --workload cfg4 --steps 2 --warmup 3 --no-cpu --no-secondary"
$C4 > $O/r02_plain_cfg4.json 2> $O/r02_plain_cfg4.err && {
  $LIST -c 400 --log-file $O/r02_launches_cfg4.csv $C4 > $O/r02_ncu_list_cfg4.log 2>&1
  $NCU -k "regex:jacobian_" -s 12 -c 2 -f -o $O/r02_cfg4 $C4 > $O/r02_ncu_cfg4.log 2>&1
}
tail -3 $O/${TAG}_pytest_gpu.txt
ls -la $O | tail -20
