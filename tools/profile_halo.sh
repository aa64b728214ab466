#!/bin/bash
# full ncu captures of the halo conv / filter-gradient kernels (template arguments print as integers)
set -u
TAG=${1:-r02}; WL=${2:-2d}
OUT=gpurun_out
export URESNET_B200_GRAPHS=0
CMD="python bench.py --workload $WL --only --steps 1 --warmup 3 --no-cpu-baseline"
for SPEC in "gconv_halo_fwd_rows|gconv_halo_kernel<.int.2, .int.0, .bool.0, .bool.1>|11" "gconv_halo_fwd|gconv_halo_kernel<.int.2, .int.0, .bool.1, .bool.1>|14" "wgrad_halo|wgrad_halo_kernel<.bool.0>|1" \
            "gconv_halo_dgrad|gconv_halo_kernel<.int.1, .int.0, .bool.0, .bool.0>|0" "gconv_halo_dgrad3|gconv_halo_kernel<.int.1, .int.0, .bool.1, .bool.0>|0" "gconv_halo_dgrad_acc|gconv_halo_kernel<.int.1, .int.1, .bool.0, .bool.0>|0"; do
  NAME=${SPEC%%|*}; REST=${SPEC#*|}; RE=${REST%%|*}; SKIP=${REST##*|}
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$RE" -s $SKIP -c 2 \
      -o $OUT/prof_${NAME}_${TAG}_$WL -f $CMD > $OUT/ncu_full_${NAME}_${TAG}_$WL.log 2>&1
  echo "full $NAME rc=$?"
  ncu -i $OUT/prof_${NAME}_${TAG}_$WL.ncu-rep --page raw --csv > $OUT/raw_${NAME}_${TAG}_$WL.csv 2>/dev/null
  python tools/ncu_raw_summary.py $OUT/raw_${NAME}_${TAG}_$WL.csv > $OUT/raw_${NAME}_${TAG}_$WL.summary.txt 2>/dev/null
  rm -f $OUT/prof_${NAME}_${TAG}_$WL.ncu-rep
done
