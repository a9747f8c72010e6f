#!/bin/bash
# Round 2, GPU call 15: what in the survivor path costs the tensor pipe 16 %?  VAR 7 = ATOMS instead of the generic atomic,
# 8 = 7 without the global store of the key (wrong results), 9 = neither atomic nor store (wrong results).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02o
mkdir -p $O
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in 7 8 9; do
  VRQ_MMA_VAR=$v PROF_ITERS=1 timeout 300 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 3 --csv --log-file $O/ncu_var$v.csv python profiles/prof_r02.py dense > $O/ncu_var$v.log 2>&1; echo "ncu var$v rc=$?"
done
VRQ_MMA_VAR=7 timeout 900 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest_var7.log 2>&1; echo "pytest var7 rc=$?"; tail -2 $O/pytest_var7.log
python - <<'PY'
import csv,glob,json
for f in sorted(glob.glob('gpurun_out/r02o/ncu_var*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
    if not hdr: print(f,'no data'); continue
    d={}
    for r in rows[hdr[0]+1:]:
        if len(r)>=15: d.setdefault(r[0],{})[r[12]]=r[14]
    for k,v in d.items():
        if float(v.get('gpu__time_duration.sum','0').replace(',',''))>5e6: print(f,k,list(v.values()))
PY
