#!/bin/bash
# A/B of the TMA producer of the images-as-rows filter-gradient kernel
set -u
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "wgrad" 2>&1 | tail -5
for V in 0 1; do
  URN_WHALO_TMA=$V timeout 300 python bench.py --workload 2d --only --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/exp_wtma_$V.err | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print('TMA=$V', d['ms_per_step'], d['value'], d.get('phases_ms_per_step'), d.get('top_launch'))
"
done
for KS in 2; do
  URN_WHALO_KS=$KS timeout 300 python bench.py --workload 2d --only --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/exp_wtma_ks$KS.err | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print('TMA KS=$KS', d['ms_per_step'], d['value'], d.get('phases_ms_per_step'), d.get('top_launch'))
"
done
URN_WHALO_YS=4 timeout 300 python bench.py --workload 2d --only --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/exp_wtma_ys4.err | python -c "
import sys, json
d = json.loads(sys.stdin.readlines()[-1])
print('TMA YS=4', d['ms_per_step'], d['value'], d.get('phases_ms_per_step'), d.get('top_launch'))
"
