#!/bin/bash
# Round 2, GPU call 21: second-generation swapped-operand kernel (bias column, 4..64 queries): parity tests, regime sweep vs the first one.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02u
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest_wide.log 2>&1; rc=$?; echo "pytest wide rc=$rc"; tail -3 $O/pytest_wide.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke.log
if [ $rc -ne 0 ]; then exit 0; fi
PROF_NQS=4,8,16,24,32,40,48,64,96,128 timeout 200 python profiles/prof_r02.py stream > $O/stream_wide.txt 2>&1; echo "stream wide rc=$?"
VRQ_MMA_WIDE=0 PROF_NQS=4,8,16,24,32,40,48,64 timeout 200 python profiles/prof_r02.py stream > $O/stream_old.txt 2>&1; echo "stream old rc=$?"
cat $O/stream_wide.txt; echo ---; cat $O/stream_old.txt
