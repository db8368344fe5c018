#!/bin/bash
# round 2, call d: tests, the driver's own two command lines with the new bench.py, ncu captures of the march kernel
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -8
( time timeout 900 python bench.py > $OUT/r02d_bench_default.json 2> $OUT/r02d_bench_default.err ) 2>&1 | grep real
tail -c 600 $OUT/r02d_bench_default.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02d_bench_default.json').read().strip().splitlines()[-1])
r=d['roofline']
print('value %.1f e2e %.1f fps %.1f timed %.2fs launches %d' % (d['value'], d['e2e']['value'], d['frames_per_s'], d['timed_region_s'], d['gpu_launches']))
print('roofline: achieved %.0f peak(L1 measured) %.0f frac %.3f nominal %.0f | l2 peak %.0f | hbm %s' % (r['achieved'], r['peak'], r['frac'], r['peak_nominal'], r['l2']['peak'], r['hbm']))
print('dense', r['dense']); print('alts', d['alternatives']); print('normals', d['normals_kernel']); print('cpu', d.get('cpu_baseline'))
PY
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02d_bench_reference.json 2> $OUT/r02d_bench_reference.err ) 2>&1 | grep real
tail -c 900 $OUT/r02d_bench_reference.json; tail -c 300 $OUT/r02d_bench_reference.err
for cfg in "march::" "march_dense::--no-ess" "march_f16::--texels f16"; do
  IFS=: read tag env flags <<< "$cfg"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/r02d_$tag \
      python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives $flags > $OUT/r02d_${tag}_ncu.log 2>&1
  tail -1 $OUT/r02d_${tag}_ncu.log | cut -c1-150
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r02d_launches.csv \
    python bench.py --steps 2 --warmup 3 --views-per-step 16 --skip-cpu-baseline --no-alternatives > $OUT/r02d_ncu_launch.log 2>&1
ls -la $OUT | tail -8
