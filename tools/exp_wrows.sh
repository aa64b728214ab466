python -m pytest tests/test_kernels_gpu.py -q -p no:cacheprovider -x -k "rows or halo_wgrad" > gpurun_out/wrows_test.log 2>&1; tail -4 gpurun_out/wrows_test.log
for v in "URN_WHALO_ROWS=0" "URN_WHALO_ROWS=1"; do
  tag=$(echo "$v" | tr ' =' '__')
  env $v python bench.py --only --no-cpu-baseline --steps 10 --per-layer gpurun_out/exp_$tag > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
done
python - <<EOF2
import json,glob
for f in sorted(glob.glob("gpurun_out/exp_URN_WHALO*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms_per_step"])
    except Exception as e: print(f,"ERR",e)
EOF2
