#!/bin/bash
# Round 2, GPU call 1: full GPU test suite, smoke, the mxf4 issue-rate microbenchmark, a first bench line with all legs,
# and the lockstep-throttle experiment (dense pass time + DRAM bytes with / without).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02a
mkdir -p $O
nvidia-smi > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt; free -g >> $O/gpu.txt
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 -x --deselect tests/test_gpu_scale.py::test_phase1_at_full_baseline_size > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 600 python -m pytest tests/test_gpu_scale.py::test_phase1_at_full_baseline_size -q > $O/pytest_100m.log 2>&1; echo "pytest100m rc=$?" | tee -a $O/pytest_100m.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/smoke.log
timeout 300 profiles/microbench/mxf4_peak > $O/mxf4_peak.txt 2>&1; echo "mxf4 rc=$?" | tee -a $O/mxf4_peak.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cfg4 > $O/bench_a.json 2> $O/bench_a.err; echo "bench_a rc=$?"
for w in 0 32 128; do
  VRQ_MMA_LOCKSTEP=$w timeout 300 python bench.py --steps 10 --no-cfg4 --no-extras --no-cpu --no-parity > $O/bench_lock$w.json 2> $O/bench_lock$w.err; echo "lock$w rc=$?"
done
for w in 0 32; do
  VRQ_MMA_LOCKSTEP=$w timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none -k regex:hamming_scan_mma_kernel -c 6 --csv --log-file $O/ncu_dense_lock$w.csv python bench.py --steps 1 --warmup 1 --no-cfg4 --no-extras --no-cpu --no-parity > $O/ncu_dense_lock$w.log 2>&1; echo "ncu lock$w rc=$?"
done
tail -3 $O/pytest.log $O/pytest_100m.log $O/smoke.log
cat $O/mxf4_peak.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02a/bench_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j['roofline']
        print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'))
    except Exception as e: print(f, 'ERR', e)
PY
