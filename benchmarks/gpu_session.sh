set -x
D=gpurun_out/r02b; mkdir -p $D
timeout 900 python -m pytest tests -m gpu -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
tail -15 $D/pytest.log
timeout 300 python benchmarks/ce_bench.py > $D/ce_c3.json 2> $D/ce.err
timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5.json 2>> $D/ce.err
cat $D/ce_c3.json $D/ce_c5.json
timeout 600 python bench.py --workload c4 --steps 100 --warmup 5 --no-also --no-cpu-baseline > $D/bench_c4.json 2> $D/bench_c4.err
python -c "import json;d=json.load(open('$D/bench_c4.json'));print('c4',d['value'],d['roofline']['frac'],d['clocks'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $D/ce_c3_launches.csv python benchmarks/ce_bench.py --iters 4 --warmup 2 --no-peaks > $D/ncu_ce.log 2>&1
