#!/bin/bash
# tc6 backward with two pairs per cluster sharing a multicast column-tile stream (CL=4) against the pair kernel (CL=2):
# parity, timing, and the TMA-only / MMA-only DBG variants (need a -DDMF_TC6_DBG build of infonce_bwd_tc6.cu)
mkdir -p gpurun_out
T=gpurun_out/r2_t12
DMF_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "supcon_bf16 or infonce_bf16_vs_fp32 or column_split or dssl_bf16 or full_size_against" > ${T}_tests.log 2>&1
echo "tests rc=$?" >> ${T}_tests.log
for v in "DMF_TC6_CL=2" "DMF_TC6_CL=4"; do
  echo "== $v" >> ${T}_kb.log
  env DMF_VERBOSE=1 $v timeout 300 python tools/kernel_bench.py --what bwd --B 65536 >> ${T}_kb.log 2>&1
  env $v timeout 300 python tools/kernel_bench.py --what bwd --B 8192 >> ${T}_kb.log 2>&1
done
touch disentagled_multimodal_fusion_b200/csrc/infonce_bwd_tc6.cu
make -C disentagled_multimodal_fusion_b200/csrc EXTRA=-DDMF_TC6_DBG > ${T}_make.log 2>&1
for v in "DMF_TC6_CL=2 DMF_TC6_DBG=7" "DMF_TC6_CL=4 DMF_TC6_DBG=7" "DMF_TC6_CL=2 DMF_TC6_DBG=1" "DMF_TC6_CL=4 DMF_TC6_DBG=1" "DMF_TC6_CL=4 DMF_TC6_DBG=6"; do
  echo "== $v" >> ${T}_kb.log
  env $v timeout 300 python tools/kernel_bench.py --what bwd --B 65536 >> ${T}_kb.log 2>&1
done
tail -5 ${T}_tests.log; cat ${T}_kb.log
