#!/bin/bash
# refresh of the C3 evidence after the lane-arrangement / pitch change (2x2 pass blocks, residues (3,2))
OUT=gpurun_out; mkdir -p $OUT; TAG=r02zz
( time timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -4 ) 2>&1 | tee $OUT/${TAG}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee $OUT/${TAG}_smoke.txt
( time timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err ) 2>&1 | grep real
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02zz_bench_default.json').read().strip().splitlines()[-1]); r=d['roofline']
print('C3: value %.1f e2e %.1f fps %.1f' % (d['value'], d['e2e']['value'], d['frames_per_s']), 'frac L1 %.3f l2 %.3f dense %.3f' % (r['frac'], r['l2']['frac'], r['dense']['frac']), d['clocks'])
PY
for cfg in "march::" "march_dense::--no-ess"; do
  IFS=: read tag env flags <<< "$cfg"
  timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_$tag \
      python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-alternatives $flags > $OUT/${TAG}_${tag}_ncu.log 2>&1
  tail -1 $OUT/${TAG}_${tag}_ncu.log | cut -c1-120
done
