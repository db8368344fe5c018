#!/bin/bash
# C5 on N GPUs: sort-last brick march with one / several samples in flight, rows / bricks
N=${1:-2}; OUT=gpurun_out; mkdir -p $OUT
run() { tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) bench.py --gpus $N --workload c5 --steps 10 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('C5 N=$N $tag: %.3f ms/frame  march slowest %.3f mean %.3f  share %.2f' % (d['ms_per_step'], d['balance']['march_ms_slowest_rank'], d['balance']['march_ms_mean_over_ranks'], d['roofline']['march_share_of_step']))" ) ; }
{
run "rows+pairs, 1 sample" PYVR_CUDA_BRICK8=0 PYVR_CUDA_TWO_SAMPLES=0 --
run "bricks, 1 sample" PYVR_CUDA_BRICK8=1 PYVR_CUDA_TWO_SAMPLES=0 --
run "bricks, 4 in flight" PYVR_CUDA_BRICK8=1 PYVR_CUDA_TWO_SAMPLES=1 --
} 2>&1 | tee $OUT/r02x_c5_N$N.txt
