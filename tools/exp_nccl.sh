N=${1:-2}
for v in "" "URESNET_B200_HIPRI=1" "URESNET_B200_DP_OVERLAP=0"; do
  echo "== [$v]"
  env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 --only 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d.get('dp_check'))"
  env $v python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 10 --warmup 3 --only --workload 3d 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('3d', round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d.get('dp_check'))"
done
