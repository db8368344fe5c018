#!/bin/bash
# round 2, call b: GPU tests of the lean kernel + A/B of launch-bound / prefetch variants on C3 (ESS and dense)
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | tail -25 > $OUT/r02b_pytest.txt; tail -12 $OUT/r02b_pytest.txt
bash tools/gpu_ab.sh "default mb8 mb6 pf2 pf4 pf2l1 mb8pf3" "-1" --no-alternatives 2>&1 | tee $OUT/r02b_ab_ess.txt
bash tools/gpu_ab.sh "default mb8 mb6 pf2 pf4 pf2l1 mb8pf3" "-1" --no-alternatives --no-ess 2>&1 | tee $OUT/r02b_ab_dense.txt
