#!/bin/bash
# GPU tests + a few bench variants (device-resident value only, no CPU baseline).
TAG=${1:-quick}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|^$" | tail -30
for variant in "" "--no-ess" "--layout linear --no-ess" "--texels f16" "--texels f16 --no-ess"; do
  name=$(echo "bench$variant" | tr -d ' -')
  timeout 600 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline $variant > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - "$OUT/${TAG}_${name}.json" "$variant" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{sys.argv[2] or 'default':32s} value={d['value']:.1f} Gs/s  fps={d['frames_per_s']:.1f}  e2e={d['e2e']['value']:.1f}  "
          f"kernel_ms/launch={r['kernel_ms_per_launch']:.2f}  fetched/ref={r['samples_fetched_per_launch']/r['samples_reference_per_launch']:.3f}  frac={r['frac']:.2f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
done
