#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02e
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scan_mma.py tests/test_gpu_index.py tests/test_gpu_multi.py -m gpu -q --maxfail=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
for t in 1 0; do
VRQ_MMA_TAIL=$t timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches_bench_tail$t.csv python bench.py --steps 2 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_launches$t.log 2>&1; echo "ncu launches rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 3 --no-cfg4 --no-extras --no-cpu --no-parity > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
PROF_NQS=64 PROF_ITERS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:few_kernel --launch-skip 1 -c 1 -o $O/ncu_mid64 python profiles/prof_r02.py stream > $O/ncu_mid.log 2>&1; echo "ncu mid rc=$?"
PROF_NQS=32,48,64,80,96 timeout 300 python profiles/prof_r02.py stream > $O/stream_mid.txt 2>&1
tail -n 4 $O/pytest.log; cat $O/stream_mid.txt
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob('gpurun_out/r02e/*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
for f in ('launches_bench_tail1.csv','launches_bench_tail0.csv'):
    rows=list(csv.reader(open('gpurun_out/r02e/'+f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    print(f)
    for r in rows[hdr+1:][-14:]:
        if len(r)>=15: print(f"{float(r[14])/1e3:10.1f} us  {r[4].split('(')[0][-50:]}  grid={r[8]}")
PY
