#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT

timeout 900 python tools/c4_shard_probe.py 2>&1 | grep -v Warning | tee $OUT/r02o_c4_shard_probe.txt
timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('normals', json.dumps(d['normals_kernel'], indent=1))" | tee $OUT/r02o_normals.txt
