#!/bin/bash
# Round 2, GPU call 27: after removing the first-generation swapped-operand kernel; tensor cores from 3 queries per pass: whole GPU suite + smoke.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02aa
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
PROF_NQS=3,5,12,33,64 timeout 200 python profiles/prof_r02.py stream > $O/stream.txt 2>&1
tail -n 3 $O/pytest.log $O/smoke.log; cat $O/stream.txt
