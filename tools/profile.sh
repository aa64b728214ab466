#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full capture of the top kernels.
# usage: tools/profile.sh <tag> [2d|3d] [kernel regexes...]
set -u
TAG=${1:-r01}; WL=${2:-2d}; shift 2 || true
KERNELS=${*:-"gconv_halo_kernel wgrad_halo_kernel gconv_tc_kernel wgrad_tc_kernel bn_bwd_apply_fast_kernel bn_bwd_reduce_fast_kernel"}
OUT=gpurun_out
export URESNET_B200_GRAPHS=0     # plain launch order (the CUDA-graph replay launches the same kernels)
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
tail -1 $OUT/plain_$TAG.log | cut -c1-300
# one step's worth of launches from the steady state (3 warm-up + 1 profiled + 3 + the timed step)
ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 640 --csv --log-file $OUT/launches_${TAG}_$WL.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 20 -c 2 -o $OUT/prof_${K}_${TAG}_$WL -f $CMD > $OUT/ncu_full_${K}_${TAG}_$WL.log 2>&1
  echo "full $K rc=$?"
  # gpurun_out is capped at 64 MiB: keep the raw-metric page as CSV, drop the report
  ncu -i $OUT/prof_${K}_${TAG}_$WL.ncu-rep --page raw --csv > $OUT/raw_${K}_${TAG}_$WL.csv 2>/dev/null
  rm -f $OUT/prof_${K}_${TAG}_$WL.ncu-rep
done
