#!/bin/bash
# session 3, call 26: last seconds of the budget -- smoke() and three parity tests on the in-tree library as committed
cd /root/repo
mkdir -p gpurun_out
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 60 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "packed or host_topk or golden_cellranger" 2>&1 | tail -1
