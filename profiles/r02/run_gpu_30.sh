#!/bin/bash
# Round 2, GPU call 30: wide kernel, A buffers handed over per K half: parity + regime sweep.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ad
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 $O/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
PROF_NQS=3,8,16,32,48,64 timeout 200 python profiles/prof_r02.py stream > $O/stream.txt 2>&1
cat $O/stream.txt
