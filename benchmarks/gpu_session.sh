set -x
D=gpurun_out/r02o; mkdir -p $D
timeout 1200 python -m pytest tests/test_pso_gpu.py -m gpu -x -q > $D/pytest.log 2>&1; echo "pytest rc=$?"
tail -n 5 $D/pytest.log
V=$PWD/eqsolver_b200/lib/variants
for rep in 1 2; do
for v in main fb tab fbtab uq3 uq3fb uq1mb3; do
  L=$V/libeqsolver_b200_$v.so; [ $v = main ] && L=$PWD/eqsolver_b200/lib/libeqsolver_b200.so
  EQS_B200_LIBRARY=$L timeout 300 python bench.py --workload c4 --steps 60 --warmup 5 --no-cpu-baseline --no-also > $D/c4_${v}_$rep.json 2>> $D/err.log
  EQS_B200_LIBRARY=$L timeout 300 python bench.py --workload c2 --steps 600 --warmup 20 --no-cpu-baseline --no-also > $D/c2_${v}_$rep.json 2>> $D/err.log
done
done
timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline > $D/c2_main_also.json 2>> $D/err.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02o/c*_*.json')):
    try:
        d=json.loads([l for l in open(f) if l.startswith('{')][-1])
        print(f.split('/')[-1], '%.4e' % d['value'], '%.4f ms' % d['ms_per_step'], 'frac %.4f' % d['roofline']['frac'], d['clocks']['sm_mhz'], d['clocks'].get('power_w'))
    except Exception as e:
        print(f, 'ERR', e)
PY
