#!/bin/bash
# One gpurun call of round 2: GPU tests, bench lines, launch list.  Outputs under gpurun_out/ (merged back).
# usage: tools/gpu_call.sh <tag> [what...]   what in: tests bench2 bench4 bench3 bench5 ncu full
tag=$1; shift
mkdir -p gpurun_out
for w in "$@"; do
  case $w in
    tests)  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" ;;
    newtests) timeout 600 python -m pytest tests/test_gpu_bench_path.py -m gpu -x -q > gpurun_out/${tag}_newtests.log 2>&1; echo "newtests rc=$?" ;;
    bench2) timeout 300 python bench.py --steps 200 --warmup 5 > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo "bench2 rc=$?" ;;
    bench4) timeout 300 python bench.py --config cfg4 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg4.json 2> gpurun_out/${tag}_bench_cfg4.err; echo "bench4 rc=$?" ;;
    bench3) timeout 420 python bench.py --config cfg3 --steps 20 --warmup 3 > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err; echo "bench3 rc=$?" ;;
    bench5) timeout 420 python bench.py --config cfg5 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg5.json 2> gpurun_out/${tag}_bench_cfg5.err; echo "bench5 rc=$?" ;;
    ref2)   timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_ref_cfg2.json 2> gpurun_out/${tag}_ref_cfg2.err; echo "ref2 rc=$?" ;;
    ncu)    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1; echo "ncu rc=$?" ;;
    smoke)  timeout 300 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?" ;;
  esac
done
tail -3 gpurun_out/${tag}_pytest.log 2>/dev/null
