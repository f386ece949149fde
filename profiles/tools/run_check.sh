python -m pytest tests -m gpu -x -q 2>&1 | tail -3; python bench.py --skip-cpu --skip-e2e 2> gpurun_out/bench_w.err | tee gpurun_out/bench_w.json | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d.get('chain_layout_resident'))"
