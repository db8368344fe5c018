#!/bin/bash
# Multi-GPU session: distributed tests + headline scaling point + C4/C5 lines.  Usage: bash tools/gpu_multi.sh N TAG
N=${1:-8}; TAG=${2:-r01m}; WHICH=${3:-all}
OUT=gpurun_out; mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -$N > $OUT/${TAG}_gpus.txt; nproc >> $OUT/${TAG}_gpus.txt; free -g | head -2 >> $OUT/${TAG}_gpus.txt
[ "$WHICH" = all ] && timeout 600 python -m pytest tests/test_sort_last_gpu.py -m gpu -q --tb=short -k across_processes 2>&1 | tail -5
run() { name=$1; shift; timeout 600 $T --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  tail -1 $OUT/${TAG}_$name.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$name: value=%.1f Gs/s fps=%.1f ms/step=%.2f e2e=%.1f march_share=%s gen=%s' % (d['value'], d['frames_per_s'], d['ms_per_step'], d['e2e']['value'], r.get('march_share_of_step'), d.get('volume_generation', d.get('normals_kernel'))))
except Exception as e:
    print('$name FAILED', e); print(open('$OUT/${TAG}_$name.err').read()[-1500:])
"; }
run c3 --steps 6 --warmup 3
run c4 --workload c4 --steps 20 --warmup 5
[ "$WHICH" = all ] && run c5_p2p --workload c5 --exchange p2p --steps 20 --warmup 5
[ "$WHICH" = all ] && run c5_nccl --workload c5 --exchange nccl --steps 20 --warmup 5
true
