#!/bin/bash
# Round 2, GPU call 19: the lean dense-pass epilogue (VAR 4) as the default: whole GPU suite, smoke, ncu tensor-pipe share, bench A/B
# against the generic loop (VRQ_MMA_VAR=0), group_tiles 8 vs 16.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02s
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 3 --csv --log-file $O/ncu_lean.csv python profiles/prof_r02.py dense > $O/ncu_lean.log 2>&1; echo "ncu rc=$?"
VRQ_MMA_GROUP_TILES=16 PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 3 --csv --log-file $O/ncu_lean_gt16.csv python profiles/prof_r02.py dense > $O/ncu_lean_gt16.log 2>&1; echo "ncu gt16 rc=$?"
for v in 4 0 4 0; do
  VRQ_MMA_VAR=$v timeout 120 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_var$v.json 2>> $O/bench_var$v.err; echo "var$v rc=$?"
done
VRQ_MMA_GROUP_TILES=16 timeout 120 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_gt16.json 2>> $O/bench_gt16.err
tail -n 3 $O/pytest.log $O/smoke.log
python - <<'PY'
import csv,glob,json
for f in sorted(glob.glob('gpurun_out/r02s/ncu_*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
    if not hdr: print(f,'no data'); continue
    d={}
    for r in rows[hdr[0]+1:]:
        if len(r)>=15: d.setdefault(r[0],{})[r[12]]=r[14]
    for k,v in d.items():
        if float(v.get('gpu__time_duration.sum','0').replace(',',''))>5e6: print(f,k,list(v.values()))
for f in sorted(glob.glob('gpurun_out/r02s/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
PY
