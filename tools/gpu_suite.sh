#!/bin/bash
# One gpurun call: GPU tests, bench variants, ncu launch list + full capture of the march kernel.
# Usage (from the repo root on the GPU box):  bash tools/gpu_suite.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt
nproc >> $OUT/${TAG}_gpu.txt
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > $OUT/${TAG}_pytest.txt
tail -15 $OUT/${TAG}_pytest.txt
summ() {
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"{sys.argv[2] or 'default':28s} value={d['value']:.1f} Gs/s  fps={d['frames_per_s']:.1f}  e2e={d['e2e']['value']:.1f}  "
          f"kernel_ms/launch={r['kernel_ms_per_launch']:.2f}  fetched/ref={r['samples_fetched_per_launch']/r['samples_reference_per_launch']:.3f}  frac={r['frac']:.2f}")
except Exception as e:
    print(sys.argv[2], "FAILED", e, open(sys.argv[1].replace('.json','.err')).read()[-800:])
PY
}
# the driver's own command lines first
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; summ $OUT/${TAG}_bench_default.json "driver default"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err; tail -c 600 $OUT/${TAG}_bench_reference.json
for variant in "--no-ess" "--layout linear --no-ess" "--texels f16" "--texels f16 --no-ess" "--texels f16 --hwtex" "--hwtex"; do
  name=$(echo "bench$variant" | tr -d ' -')
  timeout 600 python bench.py --steps 3 --warmup 3 --views-per-step 6 --skip-cpu-baseline $variant > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  summ $OUT/${TAG}_${name}.json "$variant"
done
# ncu: launch list of a short default run, then one full capture of the march kernel
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --views-per-step 12 --skip-cpu-baseline --no-alternatives > $OUT/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_march \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_march_noess \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --no-ess > $OUT/${TAG}_ncu_full_noess.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/${TAG}_march_hwtex16 \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline --texels f16 --hwtex > $OUT/${TAG}_ncu_full_hwtex16.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:normals_march -c 1 -f -o $OUT/${TAG}_normals \
    python bench.py --steps 1 --warmup 1 --views-per-step 1 --skip-cpu-baseline > $OUT/${TAG}_ncu_normals.log 2>&1
ls -la $OUT | tail -30
