#!/bin/bash
# Round 2, GPU call 3: tests (mid-regime kernel, list-free sample pass, pair scheduler), launch list of one bench step, DRAM
# bytes of the dense pass per lockstep window, full bench incl. the 1-billion-row leg.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02c
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 --deselect tests/test_gpu_scale.py::test_phase1_at_full_baseline_size > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 600 python -m pytest tests/test_gpu_scale.py::test_phase1_at_full_baseline_size -q > $O/pytest_100m.log 2>&1; echo "pytest100m rc=$?" | tee -a $O/pytest_100m.log
timeout 600 python bench.py --steps 2 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/plain.json 2> $O/plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
for w in 0 256 512; do
  VRQ_MMA_LOCKSTEP=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:hamming_scan_mma_kernel -c 6 --csv --log-file $O/ncu_dense_lock$w.csv python bench.py --steps 1 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_dense_lock$w.log 2>&1; echo "ncu lock$w rc=$?"
done
for w in 0 256 0 256; do
  VRQ_MMA_LOCKSTEP=$w timeout 300 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_lock$w.json 2>> $O/bench_lock$w.err; echo "lock$w rc=$?"
done
timeout 1500 python bench.py > $O/bench_full.json 2> $O/bench_full.err; echo "bench_full rc=$?"
timeout 300 python profiles/prof_r02.py stream > $O/stream.txt 2>&1
tail -n 4 $O/pytest.log $O/pytest_100m.log; cat $O/stream.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02c/*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln)
            r=j['roofline']
            print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
j=json.loads(open('gpurun_out/r02c/bench_full.json').read().strip().splitlines()[-1])
print(json.dumps(j.get('cfg4'),indent=1)); print({k:(round(v['ms'],3),round(v['frac'],3)) for k,v in j.items() if k.startswith('roofline_scan')})
print(json.dumps(j.get('adversarial'),indent=1))
PY
