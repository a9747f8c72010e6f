#!/bin/bash
# Round 2, GPU call 2: tests of the new kernels / limits, bench with all cfg5 variants, tail-scheduler and lockstep A/B,
# ncu --set full of the two tensor-core rescoring kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02b
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --maxfail=15 --deselect tests/test_gpu_scale.py::test_phase1_at_full_baseline_size > $O/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/pytest.log
timeout 600 python -m pytest tests/test_gpu_scale.py::test_phase1_at_full_baseline_size -q > $O/pytest_100m.log 2>&1; echo "pytest100m rc=$?" | tee -a $O/pytest_100m.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cfg4 --no-cpu > $O/bench_a.json 2> $O/bench_a.err; echo "bench_a rc=$?"
for cfg in "1 0" "0 0" "1 128" "1 512" "0 128"; do
  set -- $cfg
  VRQ_MMA_TAIL=$1 VRQ_MMA_LOCKSTEP=$2 timeout 300 python bench.py --steps 20 --no-cfg4 --no-extras --no-cpu --no-parity > $O/bench_tail$1_lock$2.json 2> $O/bench_tail$1_lock$2.err; echo "tail$1 lock$2 rc=$?"
done
PROF_PAY_ROWS=32000000 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"rescore_(binary|int8cos)_imma" -c 2 -o $O/ncu_rescore_imma python profiles/prof_r02.py rescore > $O/ncu_rescore.log 2>&1; echo "ncu rescore rc=$?"
tail -n 4 $O/pytest.log $O/pytest_100m.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02b/bench_*.json')):
    try:
        j=json.loads(open(f).read().strip().splitlines()[-1])
        r=j['roofline']
        print(f, 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'resc',round(r['rescore_ms_per_step'],3),'merge',round(r['merge_ms_per_step'],3),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'))
    except Exception as e: print(f, 'ERR', e)
j=json.loads(open('gpurun_out/r02b/bench_a.json').read().strip().splitlines()[-1])
print(json.dumps(j.get('roofline_rescore_int8cos',{}).get('imma_vs_cuda_core',{}).get('ms'),indent=1))
print(json.dumps(j.get('rescore_binary_cfg5'),indent=1))
print(json.dumps(j.get('adversarial'),indent=1))
PY
