import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tpxl_b200
from tpxl_b200 import synth
dev = "cuda:0"
sd = synth.device_state_dict(synth.dit_shapes(**synth.FULL_DIT), 1234, dev, torch.float16)
m = tpxl_b200.DiT(**synth.FULL_DIT); m.load_state_dict(sd); m = m.to(dev); del sd
x = torch.randn(1, 2048, 68, device=dev); y = torch.randn(1, 1370, 768, device=dev); t = torch.tensor([960], device=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
sync_every = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t0 = time.perf_counter()
for i in range(n):
    o = m.forward_with_cfg(x, t, y, cfg_scale=6.0, enable_amp=True)
    if (i + 1) % sync_every == 0:
        torch.cuda.synchronize()
        print("iter", i, float(o.float().abs().mean()), f"{(time.perf_counter() - t0) / (i + 1) * 1e3:.3f} ms/forward", flush=True)
print("done", flush=True)
