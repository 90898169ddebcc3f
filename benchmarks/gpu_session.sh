set -x
D=gpurun_out/r03j; mkdir -p $D
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 benchmarks/ce_bench.py --no-peaks > $D/ce_c3_n8.json 2>> $D/err.log
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29613 benchmarks/ce_bench.py --dim 1024 --samples 8388608 --objective ROSENBROCK --iters 5 --warmup 2 --no-peaks > $D/ce_c5_n8.json 2>> $D/err.log
timeout 100 python benchmarks/ce_bench.py --no-peaks > $D/ce_c3_n1.json 2>> $D/err.log
grep -h sample_evals $D/ce_*.json
