#!/bin/bash
# Round 2, GPU call 25: wide kernel with the compact survivor path (mask + one out-of-line append): parity tests + regime sweep.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02y
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest_wide.log 2>&1; rc=$?; echo "pytest wide rc=$rc"; tail -3 $O/pytest_wide.log
if [ $rc -ne 0 ]; then exit 0; fi
PROF_NQS=4,8,16,24,32,40,48,64 timeout 200 python profiles/prof_r02.py stream > $O/stream_wide.txt 2>&1; echo "stream wide rc=$?"
cat $O/stream_wide.txt
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum
PROF_NQS=64 PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:"hamming" -c 6 --csv --log-file $O/ncu_nq64.csv python profiles/prof_r02.py stream > $O/ncu_nq64.log 2>&1
grep "hamming" $O/ncu_nq64.csv | awk -F'","' '{print $1, substr($5,1,60), $13, $15}' | head -24
