O=gpurun_out; mkdir -p $O
( TPX_ATT_VARIANT=3 timeout 300 python -m pytest tests/test_gpu_attention.py -x -q -m gpu 2>&1 | tail -15
TPX_ATT_VARIANT=3 timeout 120 python tools/attn_perf.py
TPX_ATT_VARIANT=1 timeout 120 python tools/attn_perf.py
TPX_ATT_VARIANT=3 timeout 120 python tools/attn_timeline.py | tail -5 ) 2>&1 | tee $O/r02r_attn_3wg.txt
