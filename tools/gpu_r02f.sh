#!/bin/bash
# round 2, call f: full GPU test suite (TMA normals, manager shim, GUI pattern, integration stub, brick normals),
# density-first A/B (ESS + dense), bench default, ncu of the new default
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -25 | tee $OUT/r02f_pytest.txt
one() {
  local tag=$1; shift
  ( while [ "$1" != "--" ]; do export "$1"; shift; done; shift
    timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$tag: value=%.1f Gs/s ms/view=%.3f normals_ms=%s' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch'], d['normals_kernel'].get('kernel_ms')))" )
}
{
one "default ess" X=0 --
one "nodf ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_nodf.so --
one "mb8 ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_mb8.so --
one "cell4 ess" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_cell4.so --
one "default dense" X=0 -- --no-ess
one "nodf dense" PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_nodf.so -- --no-ess
one "default f16 ess" X=0 -- --texels f16
one "default f16 dense" X=0 -- --texels f16 --no-ess
one "normals no-TMA" PYVR_NORMALS_NO_TMA=1 --
} 2>&1 | tee $OUT/r02f_ab.txt
( time timeout 900 python bench.py > $OUT/r02f_bench_default.json 2> $OUT/r02f_bench_default.err ) 2>&1 | grep real
tail -c 400 $OUT/r02f_bench_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02f_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f e2e %.1f fps %.1f timed %.2fs launches %d' % (d['value'], d['e2e']['value'], d['frames_per_s'], d['timed_region_s'], d['gpu_launches']))
print('roofline: achieved %.0f peak(L1 measured) %.0f frac %.3f nominal %.0f | l2 peak %.0f | hbm %s' % (r['achieved'], r['peak'], r['frac'], r['peak_nominal'], r['l2']['peak'], r['hbm']))
print('dense', r['dense']); print('alts', d['alternatives']); print('normals', d['normals_kernel']); print('cpu', d.get('cpu_baseline'))
PY
timeout 600 python bench.py --workload c4 --steps 5 --warmup 2 > $OUT/r02f_c4_1gpu.json 2> $OUT/r02f_c4_1gpu.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02f_c4_1gpu.json').read().strip().splitlines()[-1])
    print('C4 1 GPU: %.2f ms/frame  %.1f Gsamples/s  march share %.2f' % (d['ms_per_step'], d['value'], d['roofline']['march_share_of_step']))
except Exception as e:
    print('C4 failed', e, open('gpurun_out/r02f_c4_1gpu.err').read()[-600:])
PY
for cfg in "march::" "march_dense::--no-ess"; do
  IFS=: read tag env flags <<< "$cfg"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/r02f_$tag \
      python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives $flags > $OUT/r02f_${tag}_ncu.log 2>&1
  tail -1 $OUT/r02f_${tag}_ncu.log | cut -c1-150
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^normals_" -s 2 -c 1 -f -o $OUT/r02f_normals \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives > $OUT/r02f_normals_ncu.log 2>&1
tail -1 $OUT/r02f_normals_ncu.log | cut -c1-150
