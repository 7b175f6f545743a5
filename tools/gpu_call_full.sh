#!/bin/bash
# Full evidence call: GPU tests, smoke, cfg2 bench line, reference arm, ncu launch list, ncu --set full of the hot kernels.
tag=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench2 rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_ref_cfg2.json 2> gpurun_out/${tag}_ref_cfg2.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32x3|prior_|knn_fused' -c 48 -f -o gpurun_out/${tag}_prof python tools/prof_kernels.py > gpurun_out/${tag}_prof.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/${tag}_prof.ncu-rep
timeout 300 python tools/timeline.py > gpurun_out/${tag}_timeline_cfg2.md 2> gpurun_out/${tag}_timeline.err; echo "timeline rc=$?"
timeout 300 python bench.py --config cfg4 --steps 300 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err; echo "bench4 rc=$?"
