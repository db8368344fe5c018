#!/bin/bash
# round 2, call c: where is the lean kernel bound?  ncu full captures (pair / no pair) + pair, texel and bank-rotation A/B
OUT=gpurun_out; mkdir -p $OUT
one() {  # one() tag  env...  -- bench flags
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s fps=%.1f ms/view=%.3f fetched/ref=%.3f' % (d['value'], d['frames_per_s'], r['kernel_ms_per_launch']/r['views_per_launch'], r['samples_fetched_per_launch']/r['samples_reference_per_launch']))" )
}
{
one "pair1 ess" PYVR_CUDA_PAIR=1 --
one "pair0 ess" PYVR_CUDA_PAIR=0 --
one "pair1 dense" PYVR_CUDA_PAIR=1 -- --no-ess
one "pair0 dense" PYVR_CUDA_PAIR=0 -- --no-ess
one "f16 pair1 ess" PYVR_CUDA_PAIR=1 -- --texels f16
one "f16 pair0 ess" PYVR_CUDA_PAIR=0 -- --texels f16
one "f16 pair1 dense" PYVR_CUDA_PAIR=1 -- --texels f16 --no-ess
one "linear pair1 ess" PYVR_CUDA_PAIR=1 -- --layout linear
for rx in 0 1 2 3; do for ry in 0 1 2 3; do
  one "swz $rx,$ry pair1 ess" PYVR_CUDA_PAIR=1 PYVR_CUDA_SWZ=$rx,$ry --
done; done
for s in "1,3" "1,2" "3,5" "2,1" "1,5" "3,1" "5,3" "1,7"; do
  one "swz $s pair0 ess" PYVR_CUDA_PAIR=0 PYVR_CUDA_SWZ=$s --
done
} 2>&1 | tee $OUT/r02c_ab.txt
for cfg in "march:PYVR_CUDA_PAIR=1:" "march_nopair:PYVR_CUDA_PAIR=0:" "march_dense:PYVR_CUDA_PAIR=1:--no-ess"; do
  IFS=: read tag env flags <<< "$cfg"
  env $env timeout 900 ncu --set full --clock-control none --import-source on -k "regex:march_kernel" -s 1 -c 1 -f -o $OUT/r02c_$tag \
      python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives $flags > $OUT/r02c_${tag}_ncu.log 2>&1
  tail -2 $OUT/r02c_${tag}_ncu.log | cut -c1-200
done
ls -la $OUT | tail
