set -x
D=gpurun_out/r02k; mkdir -p $D
timeout 900 python -m pytest tests/test_pso_gpu.py tests/test_random_shapes_gpu.py -m gpu -q -x > $D/pytest.log 2>&1; echo "pytest rc=$?" >> $D/pytest.log
tail -12 $D/pytest.log
for us in 2.5 1.0 4.0 8.0; do
  EQS_PSO_SEQ_ROUND_US=$us timeout 120 python benchmarks/seq_bench.py >> $D/seq.log 2>> $D/err.log
done
for w in 20000 40000 80000 160000; do
  EQS_PSO_SEQ_WINDOW=$w timeout 120 python benchmarks/seq_bench.py >> $D/seqw.log 2>> $D/err.log
done
timeout 120 python benchmarks/seq_bench.py --objective RASTRIGIN --lower -5.12 --upper 5.12 >> $D/seq_other.log 2>> $D/err.log
timeout 120 python benchmarks/seq_bench.py --objective ACKLEY --dim 256 --particles 2097152 --lower -32.768 --upper 32.768 --iters 10 >> $D/seq_other.log 2>> $D/err.log
cat $D/seq.log $D/seqw.log $D/seq_other.log | python -c "
import sys,json
for ln in sys.stdin:
    d=json.loads(ln); print(d['workload'][:50], 'sync %.4f ms  seq %.4f ms  x%.2f  seq evals/s %.3e launches/pass %.2f' % (d['synchronous']['ms_per_pass'], d['sequential_exact']['ms_per_pass'], d['sequential_over_synchronous_time'], d['sequential_exact']['evals_per_s'], d['sequential_exact']['launches_per_pass']))
"
tail -3 $D/err.log
