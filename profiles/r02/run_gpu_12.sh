#!/bin/bash
# Round 2, GPU call 12: whole GPU suite + smoke + default bench on HEAD (session-2 start), both arms.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02l
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench_full rc=$?"
tail -n 4 $O/pytest.log $O/smoke.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02l/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3), 'traffic', r['traffic'])
        except Exception as e: print(f, 'ERR', e)
PY
