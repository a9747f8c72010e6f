#!/bin/bash
# Round 2, GPU call 13: where do the dense pass's non-tensor cycles go?  Timing diagnostics of the pair kernel (results are
# wrong by construction for VAR 1..3): epilogue reads 1 of 4 column groups / reads all but filters nothing / expanders store 1 of 8.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02m
mkdir -p $O
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for v in 0 1 2 3; do
  VRQ_MMA_VAR=$v PROF_ITERS=1 timeout 300 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 6 --csv --log-file $O/ncu_var$v.csv python profiles/prof_r02.py dense > $O/ncu_var$v.log 2>&1; echo "ncu var$v rc=$?"
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r02m/ncu_var*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
    if not hdr: print(f,'no data'); continue
    d={}
    for r in rows[hdr[0]+1:]:
        if len(r)>=15: d.setdefault(r[0],{})[r[12]]=r[14]
    for k,v in d.items():
        if float(v.get('gpu__time_duration.sum','0').replace(',',''))>5e6: print(f,k,v)
PY
