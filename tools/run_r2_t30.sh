mkdir -p gpurun_out; T=gpurun_out/r2_t30
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stored" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -3 ${T}_tests.log
timeout 200 python tools/kernel_bench.py --what stored --B 65536 > ${T}_kb.log 2>&1; cat ${T}_kb.log | tail -4
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['phase_ms'], d['roofline']['avg_ms'], d['roofline']['other_kernels_ms_per_step'])"
