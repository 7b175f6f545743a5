#!/bin/bash
# multi-GPU call: tools/gpu_call_mgpu.sh <tag> <ngpu> [what...]
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1"
for w in "$@"; do
  case $w in
    check)  for m in vae hvae_2level; do timeout 300 $TR --master-port 29511 tests/manual/mgpu_check.py $m 2>&1 | grep mgpu_check >> gpurun_out/${tag}_mgpu_check.log; done; cat gpurun_out/${tag}_mgpu_check.log ;;
    weak)   timeout 300 $TR --master-port 29512 bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/${tag}_bench_n${n}_weak.json 2> gpurun_out/${tag}_bench_n${n}_weak.err; echo "weak rc=$?" ;;
    strong) timeout 300 $TR --master-port 29513 bench.py --gpus $n --steps 200 --warmup 5 --scaling strong > gpurun_out/${tag}_bench_n${n}_strong.json 2> gpurun_out/${tag}_bench_n${n}_strong.err; echo "strong rc=$?" ;;
    cfg4)   timeout 300 $TR --master-port 29514 bench.py --gpus $n --config cfg4 --steps 200 --warmup 5 > gpurun_out/${tag}_bench_n${n}_cfg4.json 2> gpurun_out/${tag}_bench_n${n}_cfg4.err; echo "cfg4 rc=$?" ;;
    sim8)   timeout 300 $TR --master-port 29515 bench.py --gpus $n --steps 200 --warmup 5 --exemplars $((3125*n)) --no-parity > gpurun_out/${tag}_bench_n${n}_sim8.json 2> gpurun_out/${tag}_bench_n${n}_sim8.err; echo "sim8 rc=$?" ;;
    bank5)  timeout 400 $TR --master-port 29517 bench.py --gpus $n --config cfg5 --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n${n}_cfg5.json 2> gpurun_out/${tag}_bench_n${n}_cfg5.err; echo "cfg5 rc=$?" ;;
    ref)    timeout 300 $TR --master-port 29516 bench.py --impl reference --gpus $n --steps 3 --warmup 1 > gpurun_out/${tag}_ref_n${n}.json 2> gpurun_out/${tag}_ref_n${n}.err; echo "ref rc=$?" ;;
  esac
done
tail -c 600 gpurun_out/${tag}_*.err 2>/dev/null
