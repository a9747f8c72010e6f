#!/bin/bash
# Round 2, GPU call 50 (4 GPUs): bench at N=4 with the final build.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02av
mkdir -p $O
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"
python - <<'PY'
import json
for ln in open('gpurun_out/r02av/bench_n4.json').read().strip().splitlines():
    if not ln.startswith('{'): continue
    j=json.loads(ln); c=j.get('cfg4',{})
    print('N=4 value',round(j['value']),'e2e',round(j['e2e']['value']),'ms',round(j['ms_per_step'],2),'parity ok',j.get('parity',{}).get('ok'),'cfg4',c.get('rows_total'),c.get('ms_per_step'),c.get('value_in_headline_unit'))
PY
