#!/bin/bash
# Same-box A/B of library builds:  profiles/ab_run.sh "head main jr3" "cfg3 cfg4"   (main = the in-tree library)
# One line per (build, workload): value, ms per pass and the dominant kernel's time -> gpurun_out/ab_<tag>.txt
LIBS=${1:-"head main"}; WLS=${2:-"cfg3 cfg4"}; TAG=${3:-ab}
mkdir -p gpurun_out
for wl in $WLS; do for lib in $LIBS; do
  unset XRSD_JR_NO_MMA
  if [ $lib = main ]; then unset XRSD_B200_LIB
  elif [ $lib = nomma ]; then unset XRSD_B200_LIB; export XRSD_JR_NO_MMA=1      # in-tree library, register-accumulator reduction
  else export XRSD_B200_LIB=$PWD/ab/libxrsd_$lib.so; fi
  python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-secondary --no-nm 2>gpurun_out/ab_err.txt | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
p = d['config']['passes_per_step']
print('%-6s %-8s value %.4g %s  ms/pass %.3f  kernel %.3f ms  frac %.3f  checksum %r' % ('$wl', '$lib', d['value'], d['unit'], d['ms_per_step'] / p, d['roofline']['kernel_ms'], d['roofline']['frac'], d['details']['checksum']))
" || tail -3 gpurun_out/ab_err.txt
done; done | tee gpurun_out/ab_$TAG.txt
