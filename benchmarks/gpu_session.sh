set -x
D=gpurun_out/r02s; mkdir -p $D
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/multigpu_check.py > $D/multigpu_p2p1.log 2>&1; echo "multigpu p2p rc=$?"
tail -n 3 $D/multigpu_p2p1.log
timeout 600 python bench.py --steps 2000 --warmup 50 --no-cpu-baseline --no-also > $D/bench_n1_2000.json 2>> $D/err.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 2000 --warmup 50 --no-cpu-baseline --no-also --no-parity > $D/bench_n2_2000.json 2>> $D/err.log
EQS_P2P=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 2000 --warmup 50 --no-cpu-baseline --no-also --no-parity > $D/bench_n2_2000_nccl.json 2>> $D/err.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $D/bench_n2_driver.json 2>> $D/err.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29545 benchmarks/seq_bench.py > $D/seq_n2.json 2>> $D/err.log
