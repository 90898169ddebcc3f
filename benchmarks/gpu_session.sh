set -x
D=gpurun_out/r02z; mkdir -p $D
timeout 1500 python -m pytest tests/test_ce_gpu.py tests/test_ce_select_gpu.py tests/test_random_shapes_gpu.py -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 $D/pytest.log
for rep in 1 2; do
timeout 300 python benchmarks/ce_bench.py --no-peaks > $D/ce_c3_$rep.json 2>> $D/err.log
timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5_$rep.json 2>> $D/err.log
done
timeout 300 python benchmarks/ce_bench.py --iters 4 --warmup 2 --no-peaks > $D/plain_c3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ce_sample_kernel -s 3 -c 1 -o $D/ce_sample_c3 python benchmarks/ce_bench.py --iters 4 --warmup 2 --no-peaks > $D/ncu_c3.log 2>&1
grep -h sample_evals $D/ce_*.json
