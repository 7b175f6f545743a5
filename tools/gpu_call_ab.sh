#!/bin/bash
# A/B call: GPU tests, cfg2 bench with the step-level switches toggled, step timeline.
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench2 rc=$?"
EXVAE_GRAPH_PRIORITY=0 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2_noprio.json 2> gpurun_out/${tag}_bench_cfg2_noprio.err; echo "bench2 (no priority) rc=$?"
EXVAE_DEFER_DW_FINISH=0 timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2_nodefer.json 2> gpurun_out/${tag}_bench_cfg2_nodefer.err; echo "bench2 (no deferred finish) rc=$?"
timeout 300 python bench.py --config cfg4 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err; echo "bench4 rc=$?"
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline_cfg2.md 2> gpurun_out/${tag}_timeline.err; echo "timeline rc=$?"
for f in gpurun_out/${tag}_bench_*.json; do python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['launches_per_step'])
"; done
