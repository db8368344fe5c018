#!/bin/bash
# Layout / texel-format sweep of the march kernel on C3 (device-resident value only).
TAG=${1:-r01b}
OUT=gpurun_out
mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | grep -v "^  \|^$" | tail -30
for variant in "--layout linear --no-ess" "--layout brick --no-ess" "--layout linear_swz --no-ess" "--layout brick_swz --no-ess" \
               "--layout linear" "--layout linear_swz" "--layout brick_swz" \
               "--layout linear --no-ess --texels f16" "--layout linear_swz --no-ess --texels f16" "--layout brick_swz --no-ess --texels f16" "--layout linear_swz --texels f16"; do
  name=$(echo "bench$variant" | tr -d ' -')
  timeout 600 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline $variant > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  python - "$OUT/${TAG}_${name}.json" "$variant" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{sys.argv[2] or 'default':48s} value={d['value']:.1f} Gs/s  fps={d['frames_per_s']:.1f}  e2e={d['e2e']['value']:.1f}  "
          f"kernel_ms/launch={r['kernel_ms_per_launch']:.2f}  fetched/ref={r['samples_fetched_per_launch']/r['samples_reference_per_launch']:.3f}  frac={r['frac']:.2f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
