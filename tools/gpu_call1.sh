#!/bin/bash
# round-2 call 1: validate the parity work on hardware
O=gpurun_out; mkdir -p $O
bash tools/run_gpu_tests.sh 500 | tee $O/r02a_tests.txt
grep -h "ref-fp16\|address reused\|strided primitives" $O/test_gpu_ref_fp16.log $O/test_gpu_sampler.log $O/test_gpu_vae.log | tee $O/r02a_ref_fp16.txt
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -6 | tee $O/r02a_smoke.txt
timeout 400 python bench.py > $O/r02a_bench.json 2> $O/r02a_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --config 5 --steps 1 --warmup 1 > $O/r02a_bench_c5.json 2> $O/r02a_bench_c5.err; echo "bench c5 rc=$?"
TPX_REF_BUDGET_S=60 timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02a_bench_ref.json 2> $O/r02a_bench_ref.err; echo "bench ref rc=$?"
nproc; lscpu | grep "Model name"
python - <<'PY'
import json
for f in ("r02a_bench","r02a_bench_c5","r02a_bench_ref"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json"))
        print(f, d["value"], d["unit"], d.get("ms_per_step"), d.get("e2e"), d.get("pipeline"), d.get("split_ms"), d.get("cpu_baseline"))
    except Exception as e:
        print(f, "ERR", e)
PY
