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
# per-phase mean durations over the steady-state iterations (fast path: tags 1 wait S, 2 S ready, 3 row loaded, 5 P buffer free, 6 exp+store done, 7 arrived)
for s, name in ((0, "WG_A"), (1, "WG_B")):
    n = int(d[s, 0])
    ev = [(int(d[s, 1 + 2 * i]), int(d[s, 2 + 2 * i]) - t0) for i in range(n)]
    iters, cur = [], {}
    for tag, t in ev:
        if tag == 1 and cur:
            iters.append(cur); cur = {}
        cur[tag] = t
    iters = [it for it in iters[2:-1] if all(k in it for k in (1, 2, 3, 5, 6, 7))]
    if not iters:
        continue
    import statistics as S
    seg = lambda a, b: S.mean(it[b] - it[a] for it in iters)  # noqa: E731
    period = S.mean(iters[i + 1][1] - iters[i][1] for i in range(len(iters) - 1))
    print(f"{name}: period {period:.0f}  wait_S {seg(1,2):.0f}  ldtm {seg(2,3):.0f}  wait_P {seg(3,5):.0f}  exp {seg(5,6):.0f}  arrive {seg(6,7):.0f}   exp starts at {[it[5] for it in iters[:4]]}")
