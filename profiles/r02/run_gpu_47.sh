#!/bin/bash
# Round 2, GPU call 47: DRAM bytes per launch of the scan regimes with the final kernels (traffic.json).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02at
mkdir -p $O
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
PROF_ITERS=1 PROF_NQS=1,2,3,16,64,96,128 timeout 300 ncu --metrics $M --clock-control none -k regex:"hamming_scan" --csv --log-file $O/traffic_stream_final.csv python profiles/prof_r02.py stream > $O/traffic_stream_final.log 2>&1; echo "rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02at/traffic_stream_final.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
d={}
for r in rows[h+1:]:
    if len(r)>=15: d.setdefault((int(r[0]),r[4][:80]),{})[r[12]]=float(r[14].replace(',',''))
for k,v in sorted(d.items()):
    if v.get('gpu__time_duration.sum',0)>1e6: print(k[0],k[1],round(v['dram__bytes_read.sum']/1e9,3),'GB read',round(v['dram__bytes_write.sum']/1e6,1),'MB written',round(v['gpu__time_duration.sum']/1e6,3),'ms')
PY
