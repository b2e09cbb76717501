# step time of the 2-GPU data-parallel training leg for several bucket plans ("bucket,first,tail" MB); 1000,0,0 = one
# bucket = the whole all-reduce exposed after backward
for cfg in 25,4,2 1000,0,0 25,0,0 8,2,1 50,8,4 25,4,8; do
  export GSD_DDP_BUCKET_MB=$cfg
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$((RANDOM%10)) bench.py --gpus 2 --steps 3 --warmup 3 --train-steps 20 --no-cpu-baseline > gpurun_out/ddp_$cfg.log 2> gpurun_out/ddp_$cfg.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/ddp_$cfg.log').read().strip().splitlines()[-1])
print('$cfg', d['train']['ddp_buckets_mb'], d['summary']['train_ms_per_step'], d['summary']['ddp_in_sync'])
PY
done
