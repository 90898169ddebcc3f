set -x
D=gpurun_out/r02c; mkdir -p $D
timeout 900 python -m pytest tests/test_ce_select_gpu.py tests/test_curand_crosscheck.py -m gpu -q > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
tail -5 $D/pytest.log
V=eqsolver_b200/lib/variants
for name in base ce_d4_mb1 ce_d4_mb3 ce_d8_mb2 philox7; do
  if [ $name = base ]; then unset EQS_B200_LIBRARY; else export EQS_B200_LIBRARY=$PWD/$V/libeqsolver_b200_$name.so; fi
  echo "== $name" >> $D/sweep.log
  timeout 300 python benchmarks/ce_bench.py --no-peaks >> $D/sweep.log 2>> $D/sweep.err
  timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks >> $D/sweep.log 2>> $D/sweep.err
done
export EQS_B200_LIBRARY=$PWD/$V/libeqsolver_b200_philox7.so
timeout 300 python bench.py --workload c4 --steps 60 --warmup 5 --no-also --no-cpu-baseline > $D/c4_philox7.json 2>> $D/sweep.err
timeout 300 python bench.py --workload c2 --steps 400 --warmup 10 --no-also --no-cpu-baseline > $D/c2_philox7.json 2>> $D/sweep.err
unset EQS_B200_LIBRARY
timeout 300 python bench.py --workload c4 --steps 60 --warmup 5 --no-also --no-cpu-baseline > $D/c4_base.json 2>> $D/sweep.err
timeout 300 python bench.py --workload c2 --steps 400 --warmup 10 --no-also --no-cpu-baseline > $D/c2_base.json 2>> $D/sweep.err
cat $D/sweep.log
for f in c4_philox7 c4_base c2_philox7 c2_base; do python -c "import json;d=json.load(open('$D/$f.json'));print('$f',d['value'],d['roofline']['frac'],d['clocks']['sm_mhz'],d['clocks']['power_w'])"; done
