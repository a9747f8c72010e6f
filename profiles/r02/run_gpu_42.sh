#!/bin/bash
# Round 2, GPU call 42 (8 GPUs): NCCL parity tests + bench at N=8 with the final build (cfg4: 1 B codes, 125 M rows per GPU).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ap
mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi_8gpu.log 2>&1; echo "pytest multi rc=$?"; tail -3 $O/pytest_multi_8gpu.log
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 10 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02ap/bench_n8.json').read().strip().splitlines():
    if not ln.startswith('{'): continue
    j=json.loads(ln); c=j.get('cfg4',{})
    print('N=8 value',round(j['value']),'e2e',round(j['e2e']['value']),'ms',round(j['ms_per_step'],2),'parity ok',j.get('parity',{}).get('ok'),'cfg4',c.get('rows_total'),c.get('ms_per_step'),c.get('value_in_headline_unit'),c.get('int8_payload','')[:40])
PY
