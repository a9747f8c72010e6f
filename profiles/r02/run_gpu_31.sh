#!/bin/bash
# Round 2, GPU call 31: sampling knobs of the headline step under the lean epilogue (bench, 30 steps each).
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02ae
mkdir -p $O
run() { n=$1; shift; env "$@" timeout 120 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_$n.json 2>> $O/bench_$n.err; echo "$n rc=$?"; }
for rep in 1 2; do
run base VRQ_X=0
run safety4 VRQ_MMA_SAFETY=4
run safety16 VRQ_MMA_SAFETY=16
run k64 VRQ_MMA_SAMPLE_K=64
run k16 VRQ_MMA_SAMPLE_K=16
run gt64 VRQ_MMA_GROUP_TILES=64
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02ae/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f.split('/')[-1], 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'scan',round(r['scan_ms_per_step'],2),'clk',j['clocks']['sm_mhz'],'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
PY
