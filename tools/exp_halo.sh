for v in "" "URN_HALO_ILV=1" "URN_HALO_ROWS=1" "URN_HALO_ROWS=1 URN_HALO_ILV=1"; do
  tag=$(echo "$v" | tr ' =' '__'); [ -z "$tag" ] && tag=base
  env $v python bench.py --only --no-cpu-baseline --steps 10 --per-layer gpurun_out/exp_$tag > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.err
  env $v python bench.py --only --no-cpu-baseline --steps 5 --workload 3d --per-layer gpurun_out/exp_$tag > gpurun_out/exp3d_$tag.json 2>> gpurun_out/exp_$tag.err
done
python - <<EOF2
import json,glob
for f in sorted(glob.glob("gpurun_out/exp*_*.json")):
    try:
        d=json.load(open(f)); print(f, round(d["value"],1), round(d["ms_per_step"],2), d["phases_ms_per_step"])
    except Exception as e: print(f,"ERR",e)
EOF2
