set -x
D=gpurun_out/r02h; mkdir -p $D
for pipe in 1 0; do
  export EQS_PSO_PIPE=$pipe
  timeout 300 python bench.py --workload c2 --steps 400 --warmup 10 --no-also --no-cpu-baseline > $D/c2_pipe$pipe.json 2>> $D/err.log
  timeout 300 python bench.py --workload c4 --steps 60 --warmup 5 --no-also --no-cpu-baseline > $D/c4_pipe$pipe.json 2>> $D/err.log
  for f in c2_pipe$pipe c4_pipe$pipe; do python -c "import json;d=json.load(open('$D/$f.json'));print('$f',d['value'],d['roofline']['frac'],d['roofline']['pattern_ceiling']['frac'],d['clocks']['sm_mhz'],d['clocks']['power_w'])"; done
done
