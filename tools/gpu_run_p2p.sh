#!/bin/bash
# mulmo_unet at N ranks: NCCL buckets (default) against the all-reduce fused into Adam over peer memory
mkdir -p gpurun_out
N=${1:-2}
for mb in 4 8; do
DNNCA_P2P_MAX_MB=$mb timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --config mulmo_unet --batch 32 --no-profile --no-cpu-baseline > gpurun_out/r02v_bench_mulmo_n${N}_p2p$mb.json 2> gpurun_out/r02v_bench_mulmo_n${N}_p2p$mb.err; echo "bench rc=$?"
tail -3 gpurun_out/r02v_bench_mulmo_n${N}_p2p$mb.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02v_bench_mulmo_n${N}_p2p$mb.json').read().strip().splitlines()[-1])
    print('p2p max MB $mb: value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e'].get('value'),'allreduce',d['config'].get('allreduce'),'loss',d.get('loss_first'),d.get('loss_last'))
except Exception as e: print('parse failed',e)
PY
done
