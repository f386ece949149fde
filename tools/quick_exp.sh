python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --skip-e2e --skip-contracts --skip-dropin --skip-cpu --skip-kernels --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d['roofline']['frac'], d['chain_layout_resident']['stage_ms'], d['nan_iv_frac'], d['gpu_launches'])"
