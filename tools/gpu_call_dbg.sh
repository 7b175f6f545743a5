#!/bin/bash
tag=$1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"
tail -6 gpurun_out/${tag}_pytest.log
if [ $rc -ne 0 ]; then
  timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_bench_path.py -m gpu -x -q -k knn_mode_step > gpurun_out/${tag}_memcheck.log 2>&1; echo "memcheck rc=$?"
  grep -E "Invalid|     at |by thread|Access at|Address|ERROR SUMMARY" gpurun_out/${tag}_memcheck.log | head -30
fi
timeout 300 python bench.py --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench2 rc=$?"
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline_cfg2.md 2> gpurun_out/${tag}_timeline.err; echo "timeline rc=$?"
for f in gpurun_out/${tag}_bench_*.json; do python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
"; done
