#!/bin/bash
# Round 2, GPU call 7: whole GPU suite + smoke on the candidate final build, lockstep windows (DRAM bytes + interleaved A/B
# timing), regime sweep, DRAM traffic of every roofline leg, full bench.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02g
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
for w in 0 64 128 256; do
  VRQ_MMA_LOCKSTEP=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_scan_mma_kernel -c 6 --csv --log-file $O/ncu_dense_lock$w.csv python bench.py --steps 1 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_dense_lock$w.log 2>&1; echo "ncu lock$w rc=$?"
done
for w in 0 128 0 128 256 0; do
  VRQ_MMA_LOCKSTEP=$w timeout 300 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_lock$w.json 2>> $O/bench_lock$w.err; echo "lock$w rc=$?"
done
timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum"
PROF_ITERS=1 timeout 600 ncu --metrics $M --clock-control none -k regex:"hamming_scan" --csv --log-file $O/traffic_stream.csv python profiles/prof_r02.py stream > $O/traffic_stream.log 2>&1; echo "traffic stream rc=$?"
PROF_ITERS=1 timeout 600 ncu --metrics $M --clock-control none -k regex:"encode" --csv --log-file $O/traffic_encode.csv python profiles/prof_r02.py encode > $O/traffic_encode.log 2>&1; echo "traffic encode rc=$?"
PROF_ITERS=1 PROF_PAY_ROWS=32000000 timeout 600 ncu --metrics $M --clock-control none -k regex:"rescore_" --csv --log-file $O/traffic_rescore.csv python profiles/prof_r02.py rescore > $O/traffic_rescore.log 2>&1; echo "traffic rescore rc=$?"
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench_full rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench_ref rc=$?"
tail -n 4 $O/pytest.log $O/smoke.log; cat $O/stream.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln)
            if 'roofline' not in j: print(f, round(j['value'],3)); continue
            r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
PY
