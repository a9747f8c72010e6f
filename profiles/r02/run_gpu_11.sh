#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02k
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_scan_mma.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.log
for e in 1 0; do
VRQ_MMA_EARLY=$e timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:hamming_scan_mma_kernel -c 6 --csv --log-file $O/ncu_dense_early$e.csv python bench.py --steps 1 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_dense_early$e.log 2>&1; echo "ncu early$e rc=$?"
done
for e in 1 0 1 0; do
  VRQ_MMA_EARLY=$e timeout 300 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_early$e.json 2>> $O/bench_early$e.err; echo "early$e rc=$?"
done
python - <<'PY'
import json,glob,csv
for f in sorted(glob.glob('gpurun_out/r02k/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
for f in sorted(glob.glob('gpurun_out/r02k/ncu_dense_early*.csv')):
    rows=list(csv.reader(open(f)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
    d={}
    for r in rows[hdr+1:]:
        if len(r)>=15: d.setdefault(r[0],{})[r[12]]=r[14]
    print(f, [v for v in d.values() if float(v.get('gpu__time_duration.sum','0'))>1e7])
PY
