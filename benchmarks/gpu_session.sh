#!/usr/bin/env bash
# One GPU-box session as used throughout round 2 (`gpurun [--gpus N] -- 'bash benchmarks/gpu_session.sh'`): the GPU parity
# suite, the driver-shaped bench lines, and -- only after the plain command has exited 0 -- the ncu captures that
# profiles/ keeps.  Everything lands under gpurun_out/<tag>/ (scratch); what is quoted anywhere is copied to profiles/.
set -x
TAG=${EQS_SESSION_TAG:-session}
D=gpurun_out/$TAG; mkdir -p $D
N=$(nvidia-smi -L | wc -l)
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > $D/smi.txt

timeout 1500 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 $D/pytest.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $D/bench_reference.json 2>> $D/err.log; echo "reference rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $D/bench_n1.json 2>> $D/err.log; echo "bench n1 rc=$?"
for n in 2 4 8; do
  [ $N -ge $n ] || continue
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n \
      tests/multigpu_check.py > $D/multigpu_n$n.log 2>&1; echo "multigpu_check n=$n rc=$?"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
      bench.py --gpus $n --steps 20 --warmup 5 > $D/bench_n$n.json 2>> $D/err.log; echo "bench n=$n rc=$?"
done

# ncu, one GPU, one launch of the top kernel of each family (the plain run of the same command first)
if [ "${EQS_SESSION_NCU:-0}" = 1 ]; then
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-also > $D/plain_c2.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pso_step_kernel -s 8 -c 1 -o $D/pso_step_c2 \
      python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-also > $D/ncu_c2.log 2>&1
  timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu-baseline --no-also > $D/plain_c4.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pso_step_kernel -s 6 -c 1 -o $D/pso_step_c4 \
      python bench.py --workload c4 --steps 8 --warmup 3 --no-cpu-baseline --no-also > $D/ncu_c4.log 2>&1
  timeout 300 python benchmarks/ce_bench.py --iters 4 --warmup 2 --no-peaks > $D/plain_c3.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:ce_sample_kernel -s 3 -c 1 -o $D/ce_sample_c3 \
      python benchmarks/ce_bench.py --iters 4 --warmup 2 --no-peaks > $D/ncu_c3.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $D/launches_bench_c2.csv \
      python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $D/ncu_launches_c2.log 2>&1
fi
ls -la $D
