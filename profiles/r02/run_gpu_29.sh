#!/bin/bash
# Round 2, GPU call 29 (2 GPUs): NCCL parity tests + bench at N=2 with the final kernels.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ac
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi_2gpu.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02ac/bench_n2.json').read().strip().splitlines():
    try:
        j=json.loads(ln); print('N=2 value',round(j['value']),'e2e',round(j['e2e']['value']),'ms',round(j['ms_per_step'],2),'parity',j.get('parity'),'cfg4',{k:j['cfg4'].get(k) for k in ('rows_total','ms_per_step','value_in_headline_unit')} if 'cfg4' in j else None)
    except Exception as e: print('ERR',e,ln[:200])
PY
