#!/bin/bash
# Longer timed regions for the partitioned lines (20 steps after 5 warm-up frames).  Usage: bash tools/gpu_multi_short.sh N TAG
N=${1:-8}; TAG=${2:-r01s}
OUT=gpurun_out; mkdir -p $OUT
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run() { name=$1; shift; timeout 300 $T --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N "$@" > $OUT/${TAG}_$name.json 2> $OUT/${TAG}_$name.err
  tail -1 $OUT/${TAG}_$name.json | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); r=d['roofline']
    print('$name: value=%.1f Gs/s fps=%.1f ms/step=%.3f e2e=%.1f march_share=%s' % (d['value'], d['frames_per_s'], d['ms_per_step'], d['e2e']['value'], r.get('march_share_of_step')))
except Exception as e:
    print('$name FAILED', e); print(open('$OUT/${TAG}_$name.err').read()[-1500:])
"; }
run c5_p2p --workload c5 --exchange p2p --steps 20 --warmup 5
run c4 --workload c4 --steps 20 --warmup 5
