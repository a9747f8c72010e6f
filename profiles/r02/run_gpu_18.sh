#!/bin/bash
# Round 2, GPU call 18: VAR 4 with the lean epilogue loop (70 instead of 151 instructions per tile on the no-survivor path).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02r
mkdir -p $O
M=gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active
VRQ_MMA_VAR=4 timeout 150 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q -x > $O/pytest_var4.log 2>&1; rc=$?; echo "pytest var4 rc=$rc"; tail -2 $O/pytest_var4.log
if [ $rc -ne 0 ]; then exit 0; fi
VRQ_MMA_VAR=4 PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 3 --csv --log-file $O/ncu_var4.csv python profiles/prof_r02.py dense > $O/ncu_var4.log 2>&1; echo "ncu var4 rc=$?"
for v in 4 0 4 0; do
  VRQ_MMA_VAR=$v timeout 120 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_var$v.json 2>> $O/bench_var$v.err; echo "var$v rc=$?"
done
python - <<'PY'
import csv,glob,json
for f in sorted(glob.glob('gpurun_out/r02r/ncu_var*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"]
    if not hdr: print(f,'no data'); continue
    d={}
    for r in rows[hdr[0]+1:]:
        if len(r)>=15: d.setdefault(r[0],{})[r[12]]=r[14]
    for k,v in d.items():
        if float(v.get('gpu__time_duration.sum','0').replace(',',''))>5e6: print(f,k,list(v.values()))
for f in sorted(glob.glob('gpurun_out/r02r/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
PY
