#!/bin/bash
# Round 2, GPU call 28: lean dense-pass epilogue on the single-CTA 128-query-tile kernel (65..128 queries, odd numbers of tiles).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ab
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 $O/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
PROF_NQS=96,128,384 timeout 200 python profiles/prof_r02.py stream > $O/stream_lean.txt 2>&1
VRQ_MMA_VAR=0 PROF_NQS=96,128,384 timeout 200 python profiles/prof_r02.py stream > $O/stream_generic.txt 2>&1
cat $O/stream_lean.txt; echo ---; cat $O/stream_generic.txt
