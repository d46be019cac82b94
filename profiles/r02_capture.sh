#!/bin/bash
# Round-2 evidence capture, run on the GPU box from the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/r02_capture.sh'
# Every ncu command is first run plainly (must exit 0) and only then under the profiler; numbers printed
# by a run under ncu are never used as bench values.  Outputs land in gpurun_out/ and are summarised
# into profiles/ afterwards (profiles/r02_summarise.py: ncu -i ... --page raw --csv).
set -u
O=gpurun_out
mkdir -p $O
FP64="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_fp64.sum"
NCU="ncu --set full --metrics $FP64 --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"

C2="python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary"
$C2 > $O/r02_plain_cfg2.json 2> $O/r02_plain_cfg2.err && {
  $LIST -c 400 --log-file $O/r02_launches_cfg2.csv $C2 > $O/r02_ncu_list_cfg2.log 2>&1
  $NCU -k regex:intensity_kernel -s 40 -c 1 -f -o $O/r02_cfg2 $C2 > $O/r02_ncu_cfg2.log 2>&1
}
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
C5="python bench.py --workload cfg5 --batch 16384 --steps 1 --warmup 1 --no-cpu --no-secondary --no-nm"
$C5 > $O/r02_plain_cfg5.json 2> $O/r02_plain_cfg5.err && {
  $LIST -c 4000 --log-file $O/r02_launches_cfg5_lm.csv $C5 > $O/r02_ncu_list_cfg5.log 2>&1
}
ls -la $O | tail -30
