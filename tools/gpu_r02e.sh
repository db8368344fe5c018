#!/bin/bash
# round 2, call e: lane arrangement x LUT layout x z-pair x bank rotation on C3 (ESS), then the driver's command lines
OUT=gpurun_out; mkdir -p $OUT
one() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" )
}
{
for lib in default arr1 arr2 lutu arr1lutu arr2lutu; do
  if [ "$lib" != "default" ]; then L="PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_$lib.so"; else L="PYVR_X=0"; fi
  for s in "3,1" "1,3" "2,3"; do one "$lib pair1 swz $s" $L PYVR_CUDA_PAIR=1 PYVR_CUDA_SWZ=$s --; done
  for s in "3,1" "3,7" "5,7"; do one "$lib pair0 swz $s" $L PYVR_CUDA_PAIR=0 PYVR_CUDA_SWZ=$s --; done
done
} 2>&1 | tee $OUT/r02e_lane_ab.txt
( time timeout 900 python bench.py > $OUT/r02e_bench_default.json 2> $OUT/r02e_bench_default.err ) 2>&1 | grep real
tail -c 400 $OUT/r02e_bench_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02e_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f e2e %.1f fps %.1f timed %.2fs launches %d' % (d['value'], d['e2e']['value'], d['frames_per_s'], d['timed_region_s'], d['gpu_launches']))
print('roofline: achieved %.0f peak(L1 measured) %.0f frac %.3f nominal %.0f | l2 peak %.0f | hbm %s' % (r['achieved'], r['peak'], r['frac'], r['peak_nominal'], r['l2']['peak'], r['hbm']))
print('dense', r['dense']); print('alts', d['alternatives']); print('normals', d['normals_kernel']); print('cpu', d.get('cpu_baseline'))
PY
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02e_bench_reference.json 2> $OUT/r02e_bench_reference.err ) 2>&1 | grep real
tail -c 1500 $OUT/r02e_bench_reference.json; tail -c 300 $OUT/r02e_bench_reference.err
