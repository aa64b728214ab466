#!/bin/bash
# compute-sanitizer on the many-tiles-per-CTA conv / filter-gradient cases (mbarrier / TMEM protocols with ring
# wrap-around).  ONE tool per gpurun call (B200_PROFILING.md).  usage: tools/sanitize.sh synccheck|racecheck|memcheck
TOOL=${1:-synccheck}
K="many_tiles and (sp0-16-16 or sp1-32-32 or sp3-128-256 or sp11 or sp12) and 1-3]"
python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider -k "$K" > gpurun_out/san_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/san_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --print-limit 20 python -m pytest tests/test_kernels_gpu.py -q -x -p no:cacheprovider -k "$K" > gpurun_out/${TOOL}_r02.log 2>&1
echo "rc=$?"; tail -12 gpurun_out/${TOOL}_r02.log; tail -2 gpurun_out/san_plain.log
