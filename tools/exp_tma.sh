python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "conv" > gpurun_out/tma_test.log 2>&1; tail -4 gpurun_out/tma_test.log
python -m pytest tests/test_net_gpu.py tests/test_parity_gpu.py -q -p no:cacheprovider -x > gpurun_out/tma_test2.log 2>&1; tail -3 gpurun_out/tma_test2.log
for v in "URN_HALO_TMA=0" "URN_HALO_TMA=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v python bench.py --only --no-cpu-baseline --steps 10 --per-layer gpurun_out/exp_$tag > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  env $v python bench.py --only --no-cpu-baseline --steps 5 --workload 3d --per-layer gpurun_out/exp_$tag > gpurun_out/exp3d_$tag.json 2>> gpurun_out/exp_$tag.err
done
python - <<EOF2
import json,glob
for f in sorted(glob.glob("gpurun_out/exp*_URN_HALO_TMA*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms_per_step"])
    except Exception as e: print(f,"ERR",e)
EOF2
for r in 1 0; do echo "=== 2D fwd tma=$r"; URN_HALO_TMA=$r python tools/halo_timeline.py 2d fwd 2>&1 | grep "per-CTA\|mma  \|epilogue:"; done
for r in 1 0; do echo "=== 3D fwd tma=$r"; URN_HALO_TMA=$r python tools/halo_timeline.py 3d fwd 2>&1 | grep "per-CTA\|mma  \|epilogue:"; done
