mkdir -p gpurun_out; T=gpurun_out/r2_t43
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm or probe or grouped_mlp or latefusion or dssl or dmvae" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -6 ${T}_tests.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['gpu_launches_per_step'], d['phase_ms'])" || tail -20 ${T}_bench_n1.log
