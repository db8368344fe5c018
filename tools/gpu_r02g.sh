#!/bin/bash
# round 2, call g (N GPUs): the multi-process GPU tests, then the driver's torchrun line with secondary + parity_check
N=${1:-2}
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_sort_last_gpu.py -m gpu -q --tb=short -k "across_processes" 2>&1 | tail -15 | tee $OUT/r02g_pytest_n$N.txt
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --steps 6 --warmup 3 > $OUT/r02g_bench_n$N.json 2> $OUT/r02g_bench_n$N.err ) 2>&1 | grep real
tail -c 1500 $OUT/r02g_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/r02g_bench_n{n}.json').read().strip().splitlines()[-1])
    print('C3 value %.1f e2e %.1f fps %.1f' % (d['value'], d['e2e']['value'], d['frames_per_s']))
    print('parity_check', json.dumps(d.get('parity_check')))
    for k, v in (d.get('secondary') or {}).items():
        if isinstance(v, dict) and 'value' in v:
            print(k, '%.2f ms/frame %.1f Gsamples/s fps %.1f march share %.2f balance %s' % (v['ms_per_step'], v['value'], v['frames_per_s'], v['roofline']['march_share_of_step'], v.get('balance')))
        else:
            print(k, v)
except Exception as e:
    print('no line', e)
PY
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 \
    bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/r02g_ref_n$N.json 2> $OUT/r02g_ref_n$N.err ) 2>&1 | grep real
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f'gpurun_out/r02g_ref_n{n}.json').read().strip().splitlines()[-1])
    print('reference arm under torchrun: value %.3f cores %s host_cores %s' % (d['value'], d['cpu_baseline']['cores'], d['cpu_baseline'].get('host_cores')))
except Exception as e:
    print('no ref line', e, open(f'gpurun_out/r02g_ref_n{n}.err').read()[-500:])
PY
