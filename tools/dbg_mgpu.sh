TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
for combo in "single 0 0" "buckets 1 0"; do
  set -- $combo
  EXVAE_GRAD_SYNC=$1 EXVAE_OVERLAP_SHARDED=$2 EXVAE_COALESCE=$3 timeout 200 $TR --master-port 29520 bench.py --gpus 2 --steps 50 --warmup 3 --no-parity > gpurun_out/dbg_$1_$2_$3.json 2> gpurun_out/dbg_$1_$2_$3.err
  echo "combo $combo rc=$?"; grep -m1 "AcceleratorError" gpurun_out/dbg_$1_$2_$3.err
done
