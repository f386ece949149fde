timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | grep -v Warning | tail -25
echo
python bench.py --skip-e2e --skip-contracts --skip-dropin --skip-cpu --skip-kernels --steps 5 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d[\"ms_per_step\"], {k:v[\"ms\"] for k,v in d[\"kernels\"].items()}, d[\"kernels_stress\"][\"missing_2pct\"], d[\"kernels_stress\"][\"shifted_ladder_1pct\"][\"stage_ms_per_day\"])"
