#!/bin/bash
# Swizzle / pair A/B on the headline workload.  Usage: bash tools/gpu_swz.sh
run() {  # pair swz flags...
  export PYVR_CUDA_PAIR=$1; if [ "$2" = "-" ]; then unset PYVR_CUDA_SWZ; else export PYVR_CUDA_SWZ=$2; fi; shift; shift
  timeout 300 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline "$@" 2>/dev/null | tail -1 | python -c "
import json,sys,os
d=json.loads(sys.stdin.read()); r=d['roofline']
print('pair=%s swz=%s %s: value=%.1f Gs/s fps=%.1f ms/launch=%.2f' % (os.environ['PYVR_CUDA_PAIR'], os.environ.get('PYVR_CUDA_SWZ','default'), ' '.join(sys.argv[1:]), d['value'], d['frames_per_s'], r['kernel_ms_per_launch']))" "$@"
}
for flags in "" "--no-ess"; do
  run 0 1,3,1 $flags; run 0 - $flags; run 0 6,4,1 $flags; run 0 2,4,3 $flags
  run 1 - $flags; run 1 1,3,1 $flags; run 1 3,2,1 $flags
done
for flags in "--texels f16" "--texels f16 --no-ess"; do
  run 0 1,3,1 $flags; run 0 - $flags
  run 1 1,3,1 $flags; run 1 - $flags
done
