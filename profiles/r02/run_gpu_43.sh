#!/bin/bash
# Round 2, GPU call 43: Phase II tensor-core kernel with a software pipeline over its candidate groups: parity + cfg5 leg.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02aq
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_index.py -m gpu -q -x > $O/pytest.log 2>&1; rc=$?; echo "pytest rc=$rc"; tail -2 $O/pytest.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python bench.py --steps 10 --no-cfg4 --no-cpu --no-parity > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.loads([l for l in open('gpurun_out/r02aq/bench.json').read().strip().splitlines() if l.startswith('{')][-1])
b=j['rescore_binary_cfg5']; r=j['roofline_rescore_int8cos']
print('phase II cfg5', b['ms'], b['frac'], b['variants_ms']); print('phase III cfg5', r['ms'], r['frac']); print('value', j['value'], 'rescore_ms_per_step', j['roofline']['rescore_ms_per_step'])
PY
