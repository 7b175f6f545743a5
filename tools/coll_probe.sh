TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
PROBE_TAG=default timeout 200 $TR --master-port 29541 tools/coll_probe.py 2>&1 | grep -E "^default|Error" | head -5
PROBE_TAG=nvls NCCL_ALGO=NVLS timeout 200 $TR --master-port 29542 tools/coll_probe.py 2>&1 | grep -E "^nvls|Error" | head -5
PROBE_TAG=tree NCCL_ALGO=Tree timeout 200 $TR --master-port 29543 tools/coll_probe.py 2>&1 | grep -E "^tree|Error" | head -5
PROBE_TAG=ll128 NCCL_PROTO=LL128 timeout 200 $TR --master-port 29544 tools/coll_probe.py 2>&1 | grep -E "^ll128|Error" | head -5
PROBE_TAG=simple NCCL_PROTO=Simple timeout 200 $TR --master-port 29545 tools/coll_probe.py 2>&1 | grep -E "^simple|Error" | head -5
