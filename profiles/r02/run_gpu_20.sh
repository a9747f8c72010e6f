#!/bin/bash
# Round 2, GPU call 20: knob sweep of the lean dense pass under the power cap (bench, 30 steps each), DRAM bytes, ncu --set full.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out/r02t
mkdir -p $O
run() { # name, env...
  n=$1; shift
  env "$@" timeout 120 python bench.py --steps 30 --no-cfg4 --no-extras --no-cpu --no-parity >> $O/bench_$n.json 2>> $O/bench_$n.err; echo "$n rc=$?"
}
for rep in 1 2; do
run base VRQ_X=0
run gt16 VRQ_MMA_GROUP_TILES=16
run gt32 VRQ_MMA_GROUP_TILES=32
run gt16_lock64 VRQ_MMA_GROUP_TILES=16 VRQ_MMA_LOCKSTEP=64
run gt16_lock128 VRQ_MMA_GROUP_TILES=16 VRQ_MMA_LOCKSTEP=128
run gt16_raw3 VRQ_MMA_GROUP_TILES=16 VRQ_MMA_RAW_STAGES=3
done
M="dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_srcunit_tex_op_read.sum"
for w in 0 64; do
VRQ_MMA_GROUP_TILES=16 VRQ_MMA_LOCKSTEP=$w PROF_ITERS=1 timeout 200 ncu --metrics $M --clock-control none -k regex:hamming_scan_mma_kernel -c 3 --csv --log-file $O/ncu_dram_lock$w.csv python profiles/prof_r02.py dense > $O/ncu_dram_lock$w.log 2>&1; echo "ncu dram lock$w rc=$?"
done
VRQ_MMA_GROUP_TILES=16 PROF_ITERS=1 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"hamming_scan_mma_kernel" --launch-skip 1 -c 1 -o $O/ncu_dense_lean python profiles/prof_r02.py dense > $O/ncu_dense_lean.log 2>&1; echo "ncu full rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02t/bench_*.json')):
    for ln in open(f).read().strip().splitlines():
        try:
            j=json.loads(ln); r=j['roofline']
            print(f.split('/')[-1], 'value',round(j['value']), 'ms',round(j['ms_per_step'],2),'dense',round(r['kernel_ms'],2),'clk',j['clocks']['sm_mhz'], j['clocks'].get('power_w_median'), 'frac', round(r['frac'],3))
        except Exception as e: print(f, 'ERR', e)
PY
grep -h "hamming_scan_mma_kernel" $O/ncu_dram_lock*.csv | awk -F'","' '{print $5, $13, $15}' | head -30
