#!/bin/bash
# Round 2, GPU call 40: same-box A/B again after removing the runtime column-group offset of the 4-warp epilogue.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02an
mkdir -p $O
OLD=$GRAFT_REPO_ROOT/vectorragquantization_b200/libvrq_ab14f51a8.so
for rep in 1 2; do
PROF_NQS=3,8,16,32,64 timeout 200 python profiles/prof_r02.py stream >> $O/new.txt 2>&1
VRQ_LIBVRQ=$OLD PROF_NQS=3,8,16,32,64 timeout 200 python profiles/prof_r02.py stream >> $O/old.txt 2>&1
done
echo new; cat $O/new.txt; echo old; cat $O/old.txt
