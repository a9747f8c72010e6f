#!/bin/bash
# Round 2, GPU call 34: swapped-operand kernel for 65..128 queries (8 epilogue warps, 4 whole-row expander warps, per K quarter, epilogue at 184 registers).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ak
mkdir -p $O
timeout 150 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -3 $O/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
PROF_NQS=64,65,80,96,112,128 timeout 150 python profiles/prof_r02.py stream > $O/stream_w128.txt 2>&1
VRQ_MMA_W128=0 PROF_NQS=65,96,128 timeout 150 python profiles/prof_r02.py stream > $O/stream_tile.txt 2>&1
cat $O/stream_w128.txt; echo ---; cat $O/stream_tile.txt
