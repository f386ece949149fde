# contracts workload A/B: the draws inside the bridge kernel vs tapes written and read (same candles)
python -m pytest tests -m gpu -x -q -k "contract or bridge or tape" 2>&1 | tail -4
for v in fused twostep; do
  extra=""; [ $v = twostep ] && extra="--ct-two-step-bridge"
  python bench.py --workload contracts --skip-cpu --skip-e2e $extra > gpurun_out/ct_$v.json 2> gpurun_out/ct_$v.err; echo "$v rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/ct_$v.json').read().strip().splitlines()[-1])
c=d.get('contracts', d)
print('$v', c['ms_per_step'], {k:(v['ms'], v['frac_hbm']) for k,v in c['stages'].items()})
PY
done
