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
for _ in range(4):
    lib.tpx_attention_tc(q.data_ptr(), k.data_ptr(), vT.data_ptr(), o.data_ptr(), B, H, Nq, Nk, Nk, Dh, Dh ** -0.5, st())
torch.cuda.synchronize()
