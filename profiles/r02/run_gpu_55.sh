#!/bin/bash
# Round 2, GPU call 55: the driver's own sequence on the final build: pytest -m gpu, smoke(), default bench, reference arm.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02aw
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
tail -n 2 $O/pytest.log $O/smoke.log
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r02aw/bench_full.json').read().strip().splitlines() if l.startswith('{')][-1]); r=j['roofline']
print('value',round(j['value']),'e2e',round(j['e2e']['value']),'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'frac',round(r['frac'],3),'launches',j['gpu_launches'],'clocks',j['clocks'])
print('cfg1',{k:(round(v['add_docs_per_s']),round(v['search_queries_per_s'])) for k,v in j['cfg1'].items() if isinstance(v,dict) and 'add_docs_per_s' in v})
PY
