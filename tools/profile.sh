#!/bin/bash
# Run on the GPU box (under gpurun): launch list of one steady-state step + full captures of the top kernels.
# usage: tools/profile.sh <tag> [2d|3d]
set -u
TAG=${1:-r02}; WL=${2:-2d}
OUT=gpurun_out
export URESNET_B200_GRAPHS=0     # plain launch order (the CUDA-graph replay launches the same kernels)
CMD="python bench.py --workload $WL --only --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_${TAG}_$WL.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain_${TAG}_$WL.log; exit 1; }
tail -1 $OUT/plain_${TAG}_$WL.log | cut -c1-200
# one step's worth of launches from the steady state (3 warm-up steps + 1 profiled step precede it)
ncu --metrics gpu__time_duration.sum --clock-control none -s 2200 -c 640 --csv --log-file $OUT/launches_${TAG}_$WL.csv $CMD > $OUT/ncu_launch_${TAG}_$WL.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py $OUT/launches_${TAG}_$WL.csv > $OUT/launches_${TAG}_$WL.summary.txt
# full captures: "<name>|<demangled-name regex>|<skip>": the full-resolution instances of each kernel family
# (the halo conv / filter-gradient kernels: tools/profile_halo.sh)
for SPEC in "bn_bwd_apply|bn_bwd_apply_fast_kernel|0" "bn_bwd_reduce|bn_bwd_reduce_fast_kernel|0" "bn_apply|bn_apply_fast_kernel|0" \
            "gconv_tc|gconv_tc_kernel|6"; do
  NAME=${SPEC%%|*}; REST=${SPEC#*|}; RE=${REST%%|*}; SKIP=${REST##*|}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RE" -s $SKIP -c 2 \
      -o $OUT/prof_${NAME}_${TAG}_$WL -f $CMD > $OUT/ncu_full_${NAME}_${TAG}_$WL.log 2>&1
  echo "full $NAME rc=$?"
  # gpurun_out is capped at 64 MiB: keep the raw-metric page as CSV, drop the report
  ncu -i $OUT/prof_${NAME}_${TAG}_$WL.ncu-rep --page raw --csv > $OUT/raw_${NAME}_${TAG}_$WL.csv 2>/dev/null
  python tools/ncu_raw_summary.py $OUT/raw_${NAME}_${TAG}_$WL.csv > $OUT/raw_${NAME}_${TAG}_$WL.summary.txt 2>/dev/null
  rm -f $OUT/prof_${NAME}_${TAG}_$WL.ncu-rep
done
