import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tpxl_b200 import _lib
from gpu_util import st
lib = _lib.lib()
B, H, Nq, Nk, Dh, DhP = 2, 16, 2048, 2048, 72, 80
q = torch.randn(B, H, Nq, DhP, device="cuda").half(); k = torch.randn(B, H, Nk, DhP, device="cuda").half()
vT = torch.randn(B, H, DhP, Nk, device="cuda").half(); o = torch.empty(B, Nq, H * Dh, device="cuda", dtype=torch.float16)
for it in range(3):
    dbg = torch.zeros(3 * 1024, dtype=torch.int64, device="cuda")
    lib.tpx_attention_tc_debug(q.data_ptr(), k.data_ptr(), vT.data_ptr(), o.data_ptr(), B, H, Nq, Nk, Nk, Dh, Dh ** -0.5, dbg.data_ptr(), st())
    torch.cuda.synchronize()
d = dbg.cpu().reshape(3, 1024)
t0 = min(int(d[s, 2]) for s in range(3) if d[s, 0] > 0)
for s, name in enumerate(("WG_A", "WG_B", "MMA")):
    n = int(d[s, 0])
    ev = [(int(d[s, 1 + 2 * i]), int(d[s, 2 + 2 * i]) - t0) for i in range(n)]
    print(name, n, "events")
    print(" ".join(f"{tag}@{t}" for tag, t in ev[:120]))
