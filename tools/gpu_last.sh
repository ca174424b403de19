#!/bin/bash
# Last GPU call of round 2 (about 2 GPU-minutes were left): the two modules that exercise the timestep table, then the bench line.
O=gpurun_out
mkdir -p $O
timeout 60 python -m pytest tests/test_gpu_sampler.py tests/test_gpu_dit.py -x -q -p no:cacheprovider 2>&1 | tail -30 > $O/r02x_tests.txt
tail -3 $O/r02x_tests.txt
timeout 45 python bench.py --no-cpu > $O/r02x_bench.json 2> $O/r02x_bench.err
echo "bench rc=$?"
python - <<'P'
import json
try:
    d = json.load(open("gpurun_out/r02x_bench.json"))
    print("steps/s", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "table", d.get("timestep_table"), "clocks", d["clocks"])
except Exception as ex:
    print("no bench line:", ex)
    print(open("gpurun_out/r02x_bench.err").read()[-1500:])
P
