#!/bin/bash
# quick call: bench-path tests, cfg2 bench (+ optional env variants), timeline
tag=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_path.py tests/test_gpu_models.py -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench2 rc=$?"
EXVAE_PREFETCH_EXEMPLARS=0 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2_noprefetch.json 2> gpurun_out/${tag}_bench_cfg2_noprefetch.err; echo "bench2 (no prefetch) rc=$?"
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline_cfg2.md 2> gpurun_out/${tag}_timeline.err; echo "timeline rc=$?"
for f in gpurun_out/${tag}_bench_*.json; do python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['launches_per_step'])
"; done
tail -3 gpurun_out/${tag}_bench_cfg2.err
