#!/bin/bash
# tc6 (M=128 pair) backward: parity + timing against the pair kernel (tc3); DBG variants need -DDMF_TC6_DBG
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "supcon_bf16 or infonce_bf16_vs_fp32 or column_split or dssl_bf16 or full_size_against" > gpurun_out/r2_t1_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2_t1_tests.log
for v in "A=1" "DMF_BWD_V3=1" "DMF_TC6_DBG=1" "DMF_TC6_DBG=2" "DMF_TC6_DBG=4" "DMF_TC6_DBG=6" "DMF_TC6_DBG=7"; do
  echo "== $v" >> gpurun_out/r2_t1_kb.log
  env $v timeout 300 python tools/kernel_bench.py --what bwd --B 65536 >> gpurun_out/r2_t1_kb.log 2>&1
done
tail -5 gpurun_out/r2_t1_tests.log; cat gpurun_out/r2_t1_kb.log
