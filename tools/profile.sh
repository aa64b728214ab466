#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full capture of the top kernels.
# usage: tools/profile.sh <tag> [2d|3d]
set -u
TAG=${1:-r01}; WL=${2:-2d}
OUT=gpurun_out
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_$TAG.log; exit 1; }
tail -1 $OUT/plain_$TAG.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file $OUT/launches_${TAG}_$WL.csv $CMD > $OUT/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gconv_tc_kernel -s 40 -c 2 -o $OUT/prof_conv_${TAG}_$WL -f $CMD > $OUT/ncu_full1_$TAG.log 2>&1
echo "full conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel -s 40 -c 2 -o $OUT/prof_wgrad_${TAG}_$WL -f $CMD > $OUT/ncu_full2_$TAG.log 2>&1
echo "full wgrad rc=$?"
ncu --set full --clock-control none --import-source on -k regex:bn_bwd_apply_kernel -s 40 -c 1 -o $OUT/prof_bnbwd_${TAG}_$WL -f $CMD > $OUT/ncu_full3_$TAG.log 2>&1
echo "full bn rc=$?"
