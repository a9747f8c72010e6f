#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02d
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scan_mma.py tests/test_gpu_classes.py -m gpu -q --maxfail=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cfg4 --no-extras --no-cpu --no-parity > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_stream.csv python profiles/prof_r02.py stream > $O/ncu_stream.log 2>&1; echo "ncu stream rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
VRQ_MMA_MID=0 timeout 300 python profiles/prof_r02.py stream > $O/stream_nomid.txt 2>&1
timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1
tail -n 4 $O/pytest.log; cat $O/stream.txt $O/stream_nomid.txt
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob('gpurun_out/r02d/*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
for f in ('launches_stream.csv','launches_bench.csv'):
    rows=list(csv.reader(open('gpurun_out/r02d/'+f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    print(f)
    for r in rows[hdr+1:][-60:]:
        if len(r)>=15: print(f"{float(r[14])/1e3:10.1f} us  {r[4].split('(')[0][-50:]}  grid={r[8]}")
PY
