#!/bin/bash
# quick GPU check of a kernel change: the IV/spline/surface GPU tests, then the resident surface bench (no CPU arm, no e2e)
TAG=${1:-q}
shift
python -m pytest tests -m gpu -x -q -k "${KEXPR:-implied or iv or surface or kats or spline or fullsize}" 2>&1 | tail -3
python bench.py --skip-cpu --skip-e2e --skip-dropin --skip-stress --skip-contracts "$@" 2> gpurun_out/bench_${TAG}.err > gpurun_out/bench_${TAG}.json
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${TAG}.json').read().strip().splitlines()[-1])
print('${TAG}', d['ms_per_step'], {k:v['ms'] for k,v in d['kernels'].items()}, d.get('chain_layout_resident',{}).get('stage_ms'), d['roofline']['frac'], d.get('checksum_close'))
PY
