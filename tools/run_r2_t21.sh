mkdir -p gpurun_out; T=gpurun_out/r2_t21
timeout 600 python -m pytest tests -x -q -m gpu -k "tc_gemm or grouped_mlp or dssl_bf16 or dmvae_bf16 or probe_heads" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --batch 8192 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_b8192.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file ${T}_launches_b8192.csv python bench.py --steps 2 --warmup 3 --graph off --batch 8192 --no-kernel-rooflines --no-loss-check --no-cpu-baseline > ${T}_ncu.log 2>&1
tail -3 ${T}_tests.log
tail -1 ${T}_bench_b8192.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['phase_ms'])"
python tools/launch_summary.py ${T}_launches_b8192.csv | head -45
