#!/bin/bash
tag=$1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_models.py tests/test_gpu_kernels.py -m gpu -q -k "conv or Conv" > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest(conv) rc=$?"
tail -4 gpurun_out/${tag}_pytest.log
timeout 200 python bench.py --config cfg3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg3.json 2> gpurun_out/${tag}_bench_cfg3.err; echo "bench3 rc=$?"
EXVAE_CONV_DW_IMPLICIT=0 timeout 200 python bench.py --config cfg3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_cfg3_im2col.json 2> gpurun_out/${tag}_bench_cfg3_im2col.err; echo "bench3 (im2col dW) rc=$?"
for f in gpurun_out/${tag}_bench_*.json; do python -c "
import json,sys
for l in open('$f'):
    if l.startswith('{'):
        d=json.loads(l); print('$f', round(d['value']), d['ms_per_step'], d['last_loss_re_kl'], {k:v for k,v in d['breakdown_ms'].items() if 'conv' in k})
"; done
