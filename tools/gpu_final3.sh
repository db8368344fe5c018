#!/bin/bash
# final refresh: unpaired f32 rows + 4^3 macrocells as defaults (lib) against 8^3 macrocells (lib_cell3)
OUT=gpurun_out; mkdir -p $OUT; TAG=r02zzz
( time timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 ) 2>&1 | tee $OUT/${TAG}_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
timeout 600 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err
PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_cell3.so timeout 600 python bench.py --skip-cpu-baseline > $OUT/${TAG}_bench_cell3.json 2> $OUT/${TAG}_bench_cell3.err
timeout 300 python bench.py --workload c4 --steps 8 --warmup 2 > $OUT/${TAG}_c4.json 2>/dev/null
PYVR_CUDA_LIB=$PWD/pyvr_b200/libpyvr_cuda_cell3.so timeout 300 python bench.py --workload c4 --steps 8 --warmup 2 > $OUT/${TAG}_c4_cell3.json 2>/dev/null
python - <<'PY'
import json
for f in ('bench_default','bench_cell3'):
    try:
        d=json.loads(open('gpurun_out/r02zzz_%s.json' % f).read().strip().splitlines()[-1]); r=d['roofline']
        print(f,'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'frac L1 %.3f l2 %.3f dense %s' % (r['frac'], r['l2']['frac'], (r.get('dense') or {}).get('frac')))
    except Exception as e: print(f,'failed',e)
for f in ('c4','c4_cell3'):
    try:
        d=json.loads(open('gpurun_out/r02zzz_%s.json' % f).read().strip().splitlines()[-1]); print(f,'%.3f ms' % d['ms_per_step'], d['volume_generation']['ms'])
    except Exception as e: print(f,'failed',e)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_march \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives > $OUT/${TAG}_march_ncu.log 2>&1
tail -1 $OUT/${TAG}_march_ncu.log | cut -c1-100
