#!/bin/bash
# Round 2, GPU call 23: per-launch times of Phase I at 32 / 48 / 64 queries per pass (wide kernel): which launch grows with N?
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02w
mkdir -p $O
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum
for nq in 32 64; do
PROF_NQS=$nq PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:"hamming|merge|verify|sample_tau|select|compact" -c 24 --csv --log-file $O/ncu_nq$nq.csv python profiles/prof_r02.py stream > $O/ncu_nq$nq.log 2>&1; echo "ncu nq$nq rc=$?"
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r02w/ncu_nq*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
    if not hdr: print(f,'no data'); continue
    d={}
    for r in rows[hdr[0]+1:]:
        if len(r)>=15: d.setdefault((r[0],r[4][:70]),{})[r[12]]=r[14]
    print(f)
    for k,v in d.items(): print('  ',k[0],k[1],list(v.values()))
PY
