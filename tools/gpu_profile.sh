#!/bin/bash
# Profiling pass on one B200 (run under gpurun): plain run, launch list, one --set full capture per named kernel.
#   tools/gpu_profile.sh <tag> "<kernel-regex> [<kernel-regex> ...]" [bench args...]
set -u
TAG=${1:-prof}; shift
KERNELS=${1:-}; shift
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-parity --hours 0.1 $*"
mkdir -p gpurun_out
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${TAG}_plain.log; exit 1; }
tail -1 gpurun_out/${TAG}_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s ${NCU_SKIP:-2} -c ${NCU_COUNT:-2} -f -o gpurun_out/${TAG}_$K $CMD > gpurun_out/${TAG}_ncu_$K.log 2>&1
  echo "full $K rc=$?"
done
