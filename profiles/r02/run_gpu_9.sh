#!/bin/bash
# Round 2, GPU call 9 (8 GPUs): bench at N=8 (weak-scaling headline + parity probe + cfg4: 1 B codes, 125 M rows per GPU with the
# int8 rows resident, and the regenerated-payload variant for the 1 -> N comparison), NCCL tests with 4 ranks.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02i
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_n8.json 2> $O/bench_n8.err; echo "bench n8 rc=$?"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus 4 --steps 10 --warmup 3 > $O/bench_n4.json 2> $O/bench_n4.err; echo "bench n4 rc=$?"
tail -n 3 $O/pytest_multi.log
python - <<'PY'
import json
for n in (8,4):
    try:
        j=json.loads(open(f'gpurun_out/r02i/bench_n{n}.json').read().strip().splitlines()[-1])
        print(n,'value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']), j['clocks'])
        print(json.dumps(j.get('parity'))); print(json.dumps(j.get('cfg4'),indent=1)); print(json.dumps(j.get('cfg4_regenerated_payload'),indent=1))
    except Exception as e: print(n,'ERR',e)
PY
tail -5 $O/bench_n8.err
