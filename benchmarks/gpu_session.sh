set -x
D=gpurun_out/r02w; mkdir -p $D
V=$PWD/eqsolver_b200/lib/variants
for rep in 1 2; do
for v in main cenodefer cemb3 ced2mb3 ced2mb4; do
  L=$V/libeqsolver_b200_$v.so; [ $v = main ] && L=$PWD/eqsolver_b200/lib/libeqsolver_b200.so
  EQS_B200_LIBRARY=$L timeout 300 python benchmarks/ce_bench.py --no-peaks > $D/c3_${v}_$rep.json 2>> $D/err.log
  EQS_B200_LIBRARY=$L timeout 300 python benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/c5_${v}_$rep.json 2>> $D/err.log
done
done
for f in $D/c*_*.json; do echo -n "$f "; python -c "
import sys,json
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('%.4f ms/pass %.4e/s' % (d['ms_per_pass'], d['sample_evals_per_s']))
"; done
