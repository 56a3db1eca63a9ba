"""Diagnostics (not a test): in-kernel phase cycle counters of the chain kernels (RN_B200_DBG=8 [+ablations])."""
import ctypes as C
import os
import sys

os.environ.setdefault("RN_B200_DBG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from relationnetworks_clevr_b200 import ops
from relationnetworks_clevr_b200._lib import lib
from tests.test_parity_gpu import _g_params


def read(tag):
    torch.cuda.synchronize()
    buf = (C.c_longlong * (160 * 16))()
    lib().rn_debug_chain_profile.argtypes = [C.c_void_p]
    assert lib().rn_debug_chain_profile(buf) == 0
    t = torch.tensor(list(buf), dtype=torch.float64).view(160, 16)[:148]
    tiles = t[:, 5].clamp(min=1) / 2      # tiles handled by slot 0 of each CTA
    names = ["gen", "wait_acc(3/tile)", "mid_epi(2/tile)", "last_epi", "total_slot0"]
    print(f"[{tag}] per slot-0 tile, mean over CTAs (cycles):")
    for i, nm in enumerate(names):
        print(f"    {nm:18s} {float((t[:, i] / tiles).mean()):10.0f}")
    print(f"    generator per tile (both slots): wait_free {float((t[:, 6] / t[:, 5].clamp(min=1)).mean()):.0f}  gen {float((t[:, 7] / t[:, 5].clamp(min=1)).mean()):.0f}")
    rounds = tiles
    print(f"    issuer per round: wait_a {float((t[:, 8] / rounds).mean()):.0f}  wait_w {float((t[:, 9] / rounds).mean()):.0f}  "
          f"issue {float((t[:, 10] / rounds).mean()):.0f}")


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    precision = sys.argv[2] if len(sys.argv) > 2 else "parity"
    n, k, Q, G, qinj = 64, 26, 128, 256, 0
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, n, k, generator=gen).cuda().requires_grad_(True)
    q = torch.randn(B, Q, generator=gen).cuda().requires_grad_(True)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen, scale=2.0):
        wb += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
    dxg = torch.randn(B, G, generator=gen).cuda()
    for it in range(2):
        xg = ops.RelationFunction.apply(x, q, qinj, precision, *wb)
        if it == 1:
            read("fwd train")
        xg.backward(dxg)
        if it == 1:
            read("dgrad")
    with torch.no_grad():
        ops.RelationFunction.apply(x.detach(), q.detach(), qinj, precision, *[t_.detach() for t_ in wb])
        read("fwd eval")


if __name__ == "__main__":
    main()
