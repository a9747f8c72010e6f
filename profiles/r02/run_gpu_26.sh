#!/bin/bash
# Round 2, GPU call 26: wide kernel default + list-free sample pass for every tensor-core regime: whole GPU suite, smoke, regime sweep, bench.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02z
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
PROF_NQS=1,2,3,4,8,16,24,32,40,48,64,96,128,256 timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1; echo "stream rc=$?"
VRQ_SCAN_MMA_MIN_NQ=3 PROF_NQS=3 timeout 100 python profiles/prof_r02.py stream > $O/stream_min3.txt 2>&1
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
tail -n 3 $O/pytest.log $O/smoke.log; cat $O/stream.txt $O/stream_min3.txt
python - <<'PY'
import json
for ln in open('gpurun_out/r02z/bench_full.json').read().strip().splitlines():
    j=json.loads(ln); r=j['roofline']
    print('value',round(j['value']),'e2e',round(j['e2e']['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'],'frac',round(r['frac'],3),'traffic',r['traffic'])
    for k in j:
        if k.startswith('roofline_scan_stream'): print(k, round(j[k]['ms'],3), round(j[k]['frac'],3))
    print('adversarial', j.get('adversarial',{}).get('ms_per_step'), 'cfg4', j.get('cfg4',{}).get('value_in_headline_unit'))
PY
