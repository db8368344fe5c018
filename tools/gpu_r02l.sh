#!/bin/bash
# round 2, call l (N GPUs): multi-process GPU tests, the driver's N-GPU bench line (C3 + secondary C4/C5 + parity_check)
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -$N > $OUT/r02l_gpus_$N.txt
timeout 900 python -m pytest tests -m gpu -q --tb=short -k "across_processes or multi_process or two_gpus or world" 2>&1 | tail -15 | tee $OUT/r02l_pytest_$N.txt
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 6 --warmup 3 > $OUT/r02l_bench_$N.json 2> $OUT/r02l_bench_$N.err ) 2>&1 | grep real
tail -c 1500 $OUT/r02l_bench_$N.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r02l_bench_$N.json').read().strip().splitlines()[-1])
    print('C3 N=$N: value %.1f e2e %.1f ms/step %.2f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
    print('secondary', json.dumps(d.get('secondary'), indent=1)[:3000])
    print('parity_check', json.dumps(d.get('parity_check'), indent=1)[:3000])
except Exception as e:
    print('bench parse failed', e)
PY
