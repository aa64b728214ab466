#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full capture of the top kernels.
# usage: tools/profile.sh <tag> [2d|3d]
set -u
TAG=${1:-r01}; WL=${2:-2d}
OUT=gpurun_out
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
tail -1 $OUT/plain_$TAG.log | cut -c1-300
# the timed step: skip the 3 warm-up steps (~650 launches each)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1950 -c 660 --csv --log-file $OUT/launches_${TAG}_$WL.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
for K in gconv_halo_kernel wgrad_halo_kernel gconv_tc_kernel wgrad_tc_kernel bn_bwd_apply_fast_kernel bn_apply_fast_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 30 -c 2 -o $OUT/prof_${K}_${TAG}_$WL -f $CMD > $OUT/ncu_full_${K}_$TAG.log 2>&1
  echo "full $K rc=$?"
done
