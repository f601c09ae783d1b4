# Sweep of the EDL kernel's CTA shape (threads, resident CTAs per SM) on the GPU box; rebuilds edl_fusion.o per point.
#   bash tools/edl_sweep.sh "128 6" "256 4" ...
set -e
cd disentagled_multimodal_fusion_b200/csrc
for cfg in "$@"; do
  set -- $cfg
  rm -f edl_fusion.o
  make EXTRA="-DEDL_THREADS=$1 -DEDL_MINB=$2" > /dev/null 2>&1
  echo "== threads $1 minblocks $2"
  (cd ../.. && python tools/kernel_bench.py --what edl | grep train)
done
