set -x
D=gpurun_out/r02u; mkdir -p $D
timeout 1500 python -m pytest tests -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 $D/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 20 --warmup 5 > $D/bench_n2_driver.json 2>> $D/err.log; echo "bench n2 rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > $D/bench_n1_driver.json 2>> $D/err.log; echo "bench n1 rc=$?"
