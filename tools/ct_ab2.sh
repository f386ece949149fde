python -m pytest tests -m gpu -x -q -k "contract or bridge or tape" 2>&1 | tail -3
for v in default gen6 gen7; do
  if [ $v = default ]; then unset IVX_LIB_PATH; else export IVX_LIB_PATH=$PWD/tools/libivx_$v.so; fi
  python bench.py --workload contracts --skip-cpu --skip-e2e > gpurun_out/ct_$v.json 2> gpurun_out/ct_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ct_$v.json').read().strip().splitlines()[-1])
c=d.get('contracts', d)
print('$v', c['ms_per_step'], {k:(v['ms'], v['frac_hbm']) for k,v in c['stages'].items()})
PY
done
