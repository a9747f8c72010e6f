#!/bin/bash
# Round 2, GPU call 24: ncu --set full of the wide kernel's dense pass at 64 queries (where do the warps wait?)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02x
mkdir -p $O
PROF_NQS=64 PROF_ITERS=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:"hamming_scan_mma_wide" --launch-skip 1 -c 1 -o $O/ncu_wide64 python profiles/prof_r02.py stream > $O/ncu_wide64.log 2>&1; echo "ncu rc=$?"
ls -la $O
