mkdir -p gpurun_out; T=gpurun_out/r2_t38
L=disentagled_multimodal_fusion_b200/libdmf_b200.so
cp $L /tmp/lib_default.so
for v in 0 1 4 5 6; do
  if [ $v = 0 ]; then cp /tmp/lib_default.so $L; else cp tools/_variants/libdmf_b200_f4v$v.so $L; fi
  echo "== forward variant $v" >> ${T}_kb.log
  timeout 120 python tools/kernel_bench.py --what fwdstore --B 65536 --iters 60 >> ${T}_kb.log 2>&1
done
cp /tmp/lib_default.so $L
cat ${T}_kb.log
