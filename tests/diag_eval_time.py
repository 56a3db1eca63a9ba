"""Diagnostics (not a test): forward-only (eval) vs training-forward time of the relation op at the bench shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationnetworks_clevr_b200 import ops
from tests.test_parity_gpu import _g_params

B, n, k, Q, G, qinj = 640, 64, 26, 128, 256, 0
gen = torch.Generator().manual_seed(0)
x = torch.randn(B, n, k, generator=gen).cuda(); q = torch.randn(B, Q, generator=gen).cuda()
wb = []
for w, b in _g_params(n, k, Q, G, qinj, gen): wb += [w.cuda(), b.cuda()]
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for prec in ("parity", "fast"):
    with torch.no_grad():
        t_eval = timeit(lambda: ops.RelationFunction.apply(x, q, qinj, prec, *wb))
    xr = x.clone().requires_grad_(True)
    t_train = timeit(lambda: ops.RelationFunction.apply(xr, q, qinj, prec, *wb))
    print(f"{prec}: eval fwd {t_eval:.3f} ms  ({B/t_eval*1e3:.0f} q/s), training fwd {t_train:.3f} ms")
