#!/bin/bash
# Round 2, GPU call 38: whole GPU suite + smoke + regime sweep + bench on the candidate final build (safety 4, swapped-operand kernel up to 96 queries).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02al
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
PROF_NQS=1,2,3,4,8,16,24,32,40,48,64,80,96,112,128,256 timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1; echo "stream rc=$?"
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
tail -n 3 $O/pytest.log $O/smoke.log; cat $O/stream.txt
python - <<'PY'
import json
for ln in open('gpurun_out/r02al/bench_full.json').read().strip().splitlines():
    j=json.loads(ln); r=j['roofline']
    print('value',round(j['value']),'e2e',round(j['e2e']['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'],'frac',round(r['frac'],3),'traffic',r['traffic'])
    for k in j:
        if k.startswith('roofline_scan_stream'): print(k, round(j[k]['ms'],3), round(j[k]['frac'],3))
    a=j.get('adversarial',{}); print('adversarial', a.get('clustered_duplicates',{}).get('ms_per_step'), a.get('forced_fallback',{}).get('ms_per_step'), 'cfg4', j.get('cfg4',{}).get('value_in_headline_unit'))
PY
