#!/bin/bash
# One ncu --set full capture (with source counters) of one workload's kernels:  profiles/ncu_one.sh cfg3 "regex" skip count tag
WL=$1; KRE=$2; SKIP=${3:-40}; CNT=${4:-1}; TAG=${5:-x}
FP64="smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,sm__inst_executed_pipe_fp64.sum"
CMD="python bench.py --workload $WL --steps 2 --warmup 3 --no-cpu --no-secondary --no-nm"
$CMD > gpurun_out/${TAG}_plain_$WL.json 2> gpurun_out/${TAG}_plain_$WL.err && \
ncu --set full --metrics $FP64 --clock-control none --import-source on -k "regex:$KRE" -s $SKIP -c $CNT -f -o gpurun_out/${TAG}_$WL $CMD > gpurun_out/${TAG}_ncu_$WL.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_$WL.log
