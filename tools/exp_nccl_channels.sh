for ch in default 2 4 8; do
  if [ "$ch" = default ]; then unset NCCL_MAX_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$((RANDOM%10)) bench.py --gpus 2 --steps 3 --warmup 3 --train-steps 20 --no-cpu-baseline > gpurun_out/nccl_$ch.log 2> gpurun_out/nccl_$ch.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/nccl_$ch.log').read().strip().splitlines()[-1])
print('$ch', d['summary']['train_ms_per_step'], d['summary']['train_samples_per_s'], d['summary']['ddp_in_sync'])
PY
done
