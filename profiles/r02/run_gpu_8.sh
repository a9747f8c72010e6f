#!/bin/bash
# Round 2, GPU call 8 (2 GPUs): NCCL tests (torchrun path and the C-ABI-only path), bench at N=2 incl. parity probe and cfg4.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02h
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_classes.py::test_cohere_float_class -m gpu -q > $O/pytest_multi.log 2>&1; echo "pytest multi rc=$?" | tee -a $O/pytest_multi.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; echo "bench n2 rc=$?"
tail -n 5 $O/pytest_multi.log
python - <<'PY'
import json
j=json.loads(open('gpurun_out/r02h/bench_n2.json').read().strip().splitlines()[-1])
print('value',round(j['value']),'ms',round(j['ms_per_step'],2),'e2e',round(j['e2e']['value']))
print(json.dumps(j.get('parity'))); print(json.dumps(j.get('cfg4'),indent=1))
PY
tail -5 $O/bench_n2.err
