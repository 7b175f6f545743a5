nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 50 > gpurun_out/clockwatch.log &
SMI=$!
timeout 100 python tools/gemm_trace.py 2>&1 | tail -3
kill $SMI
sort gpurun_out/clockwatch.log | uniq -c | sort -rn | head -12
