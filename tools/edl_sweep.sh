set -e
cd disentagled_multimodal_fusion_b200/csrc
for cfg in "256 4" "256 3" "128 6" "128 8" "512 2"; do
  set -- $cfg
  rm -f edl_fusion.o
  make EXTRA="-DEDL_THREADS=$1 -DEDL_MINB=$2" > /dev/null 2>&1
  echo "== threads $1 minblocks $2"; grep -A2 "edl_fused_kernelILi4ELi[12]ELb1" edl_fusion.ptxas.log | grep -E "Used" | cut -c1-40
  (cd ../.. && python tools/kernel_bench.py --what edl | grep train)
done
