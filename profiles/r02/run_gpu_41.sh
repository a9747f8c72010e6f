#!/bin/bash
# Round 2, GPU call 41: final build: whole GPU suite, smoke, regime sweep, both bench arms, ncu launch list of the bench command,
# ncu --set full of the dense pass (1024 queries) and of the swapped-operand dense pass at 64 queries.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ao
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --maxfail=10 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
PROF_NQS=1,2,3,4,8,16,24,32,40,48,64,80,96,112,128,256,1024 timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1; echo "stream rc=$?"
timeout 600 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_launches_bench.log 2>&1; echo "ncu launches rc=$?"
PROF_ITERS=1 timeout 400 ncu --set full --import-source on --clock-control none -k regex:"hamming_scan_mma_kernel" --launch-skip 1 -c 1 -o $O/ncu_dense_final python profiles/prof_r02.py dense > $O/ncu_dense_final.log 2>&1; echo "ncu dense rc=$?"
PROF_NQS=64 PROF_ITERS=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"hamming_scan_mma_wide" -c 1 -o $O/ncu_wide64_final python profiles/prof_r02.py stream > $O/ncu_wide64_final.log 2>&1; echo "ncu wide rc=$?"
tail -n 3 $O/pytest.log $O/smoke.log; cat $O/stream.txt
python - <<'PY'
import json
for f in ('bench_full','bench_reference'):
    for ln in open(f'gpurun_out/r02ao/{f}.json').read().strip().splitlines():
        if not ln.startswith('{'): continue
        j=json.loads(ln)
        if 'roofline' not in j: print(f, {k:j[k] for k in ('value','unit','impl','ms_per_step') if k in j}); continue
        r=j['roofline']
        print('value',round(j['value']),'e2e',round(j['e2e']['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'],'frac',round(r['frac'],3),'traffic',r['traffic'])
        for k in j:
            if k.startswith('roofline_scan_stream'): print(k, round(j[k]['ms'],3), round(j[k]['frac'],3))
PY
