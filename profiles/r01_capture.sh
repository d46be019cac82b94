#!/bin/bash
# Round-1 evidence capture, run on the GPU box from the repo root:
#   gpurun --timeout 2400 -- 'bash profiles/r01_capture.sh'
# Every ncu command is first run plainly (must exit 0) and only then under the profiler; numbers printed
# by a run under ncu are never used as bench values.  Outputs land in gpurun_out/ and are summarised
# into profiles/ afterwards (ncu -i ... --page raw --csv).
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --set full --clock-control none --import-source on"
LIST="ncu --metrics gpu__time_duration.sum --clock-control none --csv"

python bench.py > $O/bench_default.json 2> $O/bench_default.err || echo "bench default failed"
python bench.py --impl reference --steps 5 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err || echo "reference arm failed"

for W in cfg3 cfg4 cfg5; do
  python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-secondary > $O/bench_$W.json 2> $O/bench_$W.err || echo "bench $W failed"
done

C2="python bench.py --steps 3 --warmup 3 --no-cpu --no-secondary"
$C2 > /dev/null 2>&1 && {
  $LIST -c 400 --log-file $O/launches_cfg2.csv $C2 > $O/ncu_list_cfg2.log 2>&1
  $NCU -k regex:intensity_kernel -s 3 -c 1 -f -o $O/r01c_cfg2 $C2 > $O/ncu_cfg2.log 2>&1
}
C3="python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu --no-secondary"
$C3 > /dev/null 2>&1 && {
  $LIST -c 400 --log-file $O/launches_cfg3.csv $C3 > $O/ncu_list_cfg3.log 2>&1
  $NCU -k "regex:intensity_kernel|reflection_table" -s 10 -c 2 -f -o $O/r01c_cfg3 $C3 > $O/ncu_cfg3.log 2>&1
}
C4="python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu --no-secondary"
$C4 > /dev/null 2>&1 && {
  $LIST -c 400 --log-file $O/launches_cfg4.csv $C4 > $O/ncu_list_cfg4.log 2>&1
  $NCU -k "regex:jacobian_|reflection_table" -s 13 -c 3 -f -o $O/r01c_cfg4 $C4 > $O/ncu_cfg4.log 2>&1
}
C5="python bench.py --steps 3 --warmup 3 --no-cpu"
$C5 > /dev/null 2>&1 && {
  $NCU -k regex:feature_kernel -s 1 -c 1 -f -o $O/r01c_features $C5 > $O/ncu_features.log 2>&1
}
ls -la $O | tail -30
