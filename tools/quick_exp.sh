timeout 900 python -m pytest tests -m gpu -x -q -k "ohlcv or candle or reconstruct or contract" 2>&1 | tail -3
python bench.py --skip-e2e --skip-contracts --skip-dropin --skip-cpu --skip-stress --n-t 600 --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(v['ms'],v['frac_hbm']) for k,v in d['kernels'].items()})"
