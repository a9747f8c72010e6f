#!/bin/bash
# Round 2, GPU call 10: ncu --set full of the dense pass (final scheduler) and of the two tensor-core rescoring kernels (final shapes).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02j
mkdir -p $O
PROF_ITERS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"hamming_scan_mma_kernel" --launch-skip 1 -c 1 -o $O/ncu_dense python profiles/prof_r02.py dense > $O/ncu_dense.log 2>&1; echo "ncu dense rc=$?"
PROF_ITERS=1 PROF_PAY_ROWS=32000000 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rescore_(binary|int8cos)_imma" -c 2 -o $O/ncu_rescore_imma python profiles/prof_r02.py rescore > $O/ncu_rescore.log 2>&1; echo "ncu rescore rc=$?"
timeout 600 python -m pytest tests/test_gpu_classes.py -m gpu -q > $O/pytest_classes.log 2>&1; echo "pytest classes rc=$?"; tail -3 $O/pytest_classes.log
cat $O/ncu_dense.log | tail -3
