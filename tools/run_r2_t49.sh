mkdir -p gpurun_out; T=gpurun_out/r2_t49
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "stored or rowcol or unit_norm or full_size_against or supcon or infonce" > ${T}_tests.log 2>&1; echo "tests rc=$?" >> ${T}_tests.log
tail -6 ${T}_tests.log
timeout 120 python tools/kernel_bench.py --what fwdstore --B 65536 --iters 60 > ${T}_kb.log 2>&1; tail -2 ${T}_kb.log
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline --no-kernel-rooflines > ${T}_bench_n1.log 2>&1
tail -1 ${T}_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'], d['loss_check']['loss'], d['loss_check']['grad_norm'], d['phase_ms']); print(d['roofline_k2_fwd']['frac'], d['roofline_k2_bwd']['frac'])"
