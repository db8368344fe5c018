#!/bin/bash
# round 2, call p: centre-out row order: parity subset, C4 shard probe again, C3 and C4 lines
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests/test_render_gpu.py tests/test_sort_last_gpu.py -m gpu -q --tb=short -x 2>&1 | tail -5 | tee $OUT/r02p_pytest.txt
timeout 900 python tools/c4_shard_probe.py 2>&1 | grep -v Warning | tee $OUT/r02p_c4_shard_probe.txt
timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 8 --skip-cpu-baseline --no-alternatives 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; n=d['normals_kernel']
print('C3: value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))
print('normals: %.4f ms frac hbm %.3f mix peak %.0f frac mix %.3f relaxed %.4f ms' % (n['kernel_ms'], n['frac_of_hbm_peak'], n['mix_1r3w_peak_GB/s'], n['frac_of_mix_peak'], n['relaxed_opt_in']['kernel_ms']))" | tee $OUT/r02p_c3.txt
timeout 300 python bench.py --steps 2 --warmup 2 --views-per-step 1 --skip-cpu-baseline --no-alternatives 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('C3 one view per launch: value=%.1f Gs/s ms/view=%.3f' % (d['value'], r['kernel_ms_per_launch']/r['views_per_launch']))" | tee -a $OUT/r02p_c3.txt
