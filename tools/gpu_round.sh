#!/bin/bash
# One gpurun call that produces everything profiles/ needs for a round (every step under its own timeout):
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r02'      then here:  python tools/ncu_summary.py r02
# Steps: GPU parity suite (one process per module), smoke, bench line, micro-benchmarks, attention offset sweep,
# ncu launch list, ncu --set full of the GEMM and attention kernels of one DiT block.  Outputs under gpurun_out/.
TAG=${1:-rXX}
O=gpurun_out
mkdir -p $O
bash tools/run_gpu_tests.sh 300 | tee $O/${TAG}_tests.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $O/${TAG}_smoke.txt
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"
timeout 200 python tools/gemm_perf.py 2>&1 | grep -v "tile=-" > $O/${TAG}_microbench.txt; echo "microbench rc=$?"
for sg in 0 1200 1400 1600 1800 2000; do TPX_ATT_STAGGER=$sg timeout 100 python tools/attn_perf.py 2>&1 | tail -1; done | tee $O/${TAG}_attn_sweep.txt
timeout 100 python tools/primsdf_perf.py 2>&1 | tail -4 | tee $O/${TAG}_primsdf.txt
timeout 100 python tools/raymarch_perf.py 2>&1 | tail -1 | tee $O/${TAG}_raymarch.txt
timeout 100 python tools/vae_decode_perf.py 2>&1 | tail -2 | tee $O/${TAG}_vae_decode.txt
timeout 100 python tools/vae_perf.py > $O/${TAG}_vae_kernels.txt 2>&1
timeout 100 python tools/gemm_timeline.py > $O/${TAG}_gemm_timeline.txt 2>&1
timeout 200 python bench.py --samples-per-gpu 4 --steps 10 --no-vae --no-cpu > $O/${TAG}_bench_b4.json 2> /dev/null; echo "bench B=4 rc=$?"
timeout 300 python bench.py --config 5 --steps 1 --warmup 1 > $O/${TAG}_bench_c5.json 2> /dev/null; echo "bench config5 rc=$?"
timeout 280 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:"gemm_tc|attention|ln_modulate|gemv|x_embed|cfg_combine|ddim_step" \
    -s 1100 -c 700 --csv --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-vae --no-cpu > $O/${TAG}_ncu_l.log 2>&1; echo "launch list rc=$?"
timeout 250 ncu --set full --clock-control none --import-source on -k regex:attention_tc -s 60 -c 2 -f -o $O/${TAG}_prof_attn \
    python bench.py --steps 1 --warmup 3 --no-vae --no-cpu > $O/${TAG}_ncu_a.log 2>&1; echo "attention capture rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 300 -c 6 -f -o $O/${TAG}_prof_gemm \
    python bench.py --steps 1 --warmup 3 --no-vae --no-cpu > $O/${TAG}_ncu_g.log 2>&1; echo "gemm capture rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"conv3_halo|groupnorm" -s 9 -c 6 -f -o $O/${TAG}_prof_vae \
    python tools/vae_decode_perf.py 1 > $O/${TAG}_ncu_v.log 2>&1; echo "vae capture rc=$?"
python -c "import json; d=json.load(open('$O/${TAG}_bench.json')); print('steps/s', round(d['value'],1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value'],1), 'clocks', d['clocks'])"
