mkdir -p gpurun_out; T=gpurun_out/r2_t53
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tc_gemm or grouped_mlp or dssl_bf16 or probe_heads or dmvae_bf16 or cast_dual" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -4 ${T}_tests.log
timeout 200 python tools/kernel_bench.py --what gemm --B 65536 > ${T}_kb.log 2>&1; grep -i 'gemm\|TFLOP' ${T}_kb.log | tail -8
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],2), {k: round(v,2) for k,v in d['phase_ms'].items() if k in ('mlp_fwd','mlp_bwd','rowlse_x4','infonce_bwd','graph_total')})"
