timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
for M in staged direct; do
  PG_REMOTE_SCATTER=$M timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 2 --warmup 2 --rows 134217728 --e2e-rows 1048576 --extra-scale 1.0 --only partitioned_output_all_to_all,q3_repartitioned 2>&1 | tail -1 > /tmp/o.json
  python -c "
import json
d=json.load(open('/tmp/o.json')); e=d['extra']
print('$M', e['partitioned_output_all_to_all']['fused_scatter_exchange_ms'], e['partitioned_output_all_to_all']['fused_gbs_per_gpu'], e['q3_repartitioned']['ms'])
"
done
