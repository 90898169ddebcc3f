set -x
D=gpurun_out/r02m; mkdir -p $D
# own bounds checks: the variant library under the smoke script and the select / pso tests
EQS_B200_LIBRARY=$PWD/eqsolver_b200/lib/variants/libeqsolver_b200_bounds.so timeout 600 python tests/sanitizer_smoke.py > $D/bounds_smoke.log 2>&1; echo "bounds smoke rc=$?"
EQS_B200_LIBRARY=$PWD/eqsolver_b200/lib/variants/libeqsolver_b200_bounds.so timeout 900 python -m pytest tests/test_ce_select_gpu.py tests/test_ce_gpu.py tests/test_pso_gpu.py tests/test_random_shapes_gpu.py -m gpu -q -x -k "not user_objective_hook and not out_of_device_memory" > $D/bounds_pytest.log 2>&1; echo "bounds pytest rc=$?"
tail -3 $D/bounds_smoke.log; tail -4 $D/bounds_pytest.log
timeout 600 python -m pytest tests/test_cpp_and_rust_bindings.py -m gpu -q > $D/cpp.log 2>&1; tail -3 $D/cpp.log
# CrossEntropy: one GPU vs two (p2p mailboxes vs NCCL between the phases)
timeout 300 python benchmarks/ce_bench.py --no-peaks > $D/ce_c3_n1.json 2>> $D/err.log
timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5_n1.json 2>> $D/err.log
for p2p in 1 0; do
EQS_P2P=$p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2961$p2p benchmarks/ce_bench.py --no-peaks > $D/ce_c3_n2_p2p$p2p.json 2>> $D/err.log
EQS_P2P=$p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$p2p benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5_n2_p2p$p2p.json 2>> $D/err.log
done
grep -h sample_evals $D/ce_*.json | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print(d['n_gpus'], d['workload'][:40], '%.4f ms/pass %.4e/s launches %.1f' % (d['ms_per_pass'], d['sample_evals_per_s'], d['kernel_launches_per_pass']))
"
ls $D
