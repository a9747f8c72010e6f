#!/bin/bash
# Round 2, GPU call 39: same-box A/B of the regime sweep: current build vs the build of commit 14f51a8 (before safety 4 / the 65..96 variant).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02am
mkdir -p $O
OLD=$GRAFT_REPO_ROOT/vectorragquantization_b200/libvrq_ab14f51a8.so
for rep in 1 2; do
PROF_NQS=3,16,32,64,256 timeout 200 python profiles/prof_r02.py stream >> $O/new.txt 2>&1
VRQ_LIBVRQ=$OLD PROF_NQS=3,16,32,64,256 timeout 200 python profiles/prof_r02.py stream >> $O/old.txt 2>&1
VRQ_MMA_SAFETY=8 PROF_NQS=3,16,32,64,256 timeout 200 python profiles/prof_r02.py stream >> $O/new_safety8.txt 2>&1
done
echo new; cat $O/new.txt; echo old; cat $O/old.txt; echo new_safety8; cat $O/new_safety8.txt
