#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
( time timeout 1500 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -6 ) 2>&1 | tee $OUT/r02v_pytest.txt
timeout 600 python bench.py --workload c4 --steps 8 --warmup 2 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 1 GPU: %.2f ms/frame %.1f Gs/s layout=%s' % (d['ms_per_step'], d['value'], d['config'].get('texel_layout')))" | tee $OUT/r02v_c4.txt
timeout 600 python bench.py --workload c4 --steps 8 --warmup 2 --hwtex 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C4 1 GPU hwtex: %.2f ms/frame %.1f Gs/s' % (d['ms_per_step'], d['value']))" | tee -a $OUT/r02v_c4.txt
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:^march_kernel" -s 1 -c 1 -f -o $OUT/r02v_c4_multi \
    python bench.py --workload c4 --steps 1 --warmup 1 > $OUT/r02v_c4_multi_ncu.log 2>&1
tail -1 $OUT/r02v_c4_multi_ncu.log | cut -c1-150
