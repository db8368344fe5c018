#!/bin/bash
# Final single-GPU evidence of round 2, exactly what the driver runs plus the profiles DESIGN.md cites.
OUT=gpurun_out; mkdir -p $OUT; TAG=r02z
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 ) 2>&1 | tee $OUT/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
( time timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err ) 2>&1 | grep real
tail -c 300 $OUT/${TAG}_bench_default.err
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ) 2>&1 | grep real
tail -c 300 $OUT/${TAG}_bench_reference.err
timeout 600 python bench.py --workload c4 --steps 8 --warmup 2 > $OUT/${TAG}_c4_1gpu.json 2> $OUT/${TAG}_c4_1gpu.err
timeout 600 python bench.py --workload c4 --steps 8 --warmup 2 --hwtex > $OUT/${TAG}_c4_1gpu_hwtex.json 2> $OUT/${TAG}_c4_1gpu_hwtex.err
python - <<'PY'
import json
def last(p):
    try: return json.loads(open(p).read().strip().splitlines()[-1])
    except Exception as e: return {"error": str(e)}
d=last('gpurun_out/r02z_bench_default.json'); r=d.get('roofline',{})
print('C3: value %.1f e2e %.1f fps %.1f timed %.2fs launches %s' % (d.get('value',0), d.get('e2e',{}).get('value',0), d.get('frames_per_s',0), d.get('timed_region_s',0), d.get('gpu_launches')))
print('roofline frac(L1) %.3f l2 %.3f hbm %s dense %s' % (r.get('frac',0), r.get('l2',{}).get('frac',0), (r.get('hbm') or {}).get('frac'), (r.get('dense') or {}).get('frac')))
print('normals', d.get('normals_kernel')); print('alts', d.get('alternatives')); print('cpu', d.get('cpu_baseline'))
ref=last('gpurun_out/r02z_bench_reference.json'); print('reference arm', {k: ref.get(k) for k in ('value','unit','impl','ms_per_step','cpu_baseline')})
for f in ('c4_1gpu','c4_1gpu_hwtex'):
    x=last('gpurun_out/r02z_%s.json' % f); print(f, x.get('ms_per_step'), x.get('value'))
PY
# launch list of the default bench command (short), then one --set full capture per kernel of interest
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --views-per-step 12 --skip-cpu-baseline --no-alternatives > $OUT/${TAG}_launches.log 2>&1
for cfg in "march::" "march_dense::--no-ess" "march_f16::--texels f16"; do
  IFS=: read tag env flags <<< "$cfg"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_$tag \
      python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives $flags > $OUT/${TAG}_${tag}_ncu.log 2>&1
  tail -1 $OUT/${TAG}_${tag}_ncu.log | cut -c1-150
done
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^normals_" -s 2 -c 1 -f -o $OUT/${TAG}_normals \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives > $OUT/${TAG}_normals_ncu.log 2>&1
tail -1 $OUT/${TAG}_normals_ncu.log | cut -c1-150
