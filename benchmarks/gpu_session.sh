set -x
D=gpurun_out/r03c; mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 4 $D/pytest.log
for rep in 1 2; do
timeout 300 python benchmarks/ce_bench.py --no-peaks > $D/ce_c3_$rep.json 2>> $D/err.log
timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5_$rep.json 2>> $D/err.log
timeout 300 python benchmarks/ce_bench.py --objective ACKLEY --no-peaks > $D/ce_c3ackley_$rep.json 2>> $D/err.log
done
grep -h sample_evals $D/ce_*.json
