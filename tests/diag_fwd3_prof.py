"""Diagnostics (not a test): in-kernel cycle counters of the 3-pass training forward (RN_B200_DBG=8)."""
import ctypes as C
import os
import sys

os.environ.setdefault("RN_B200_DBG", "8")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from relationnetworks_clevr_b200 import ops
from relationnetworks_clevr_b200._lib import lib
from tests.test_parity_gpu import _g_params


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    n, k, Q, G, qinj = 64, 26, 128, 256, 0
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, n, k, generator=gen).cuda().requires_grad_(True)
    q = torch.randn(B, Q, generator=gen).cuda().requires_grad_(True)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen, scale=2.0):
        wb += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
    for _ in range(2):
        ops.RelationFunction.apply(x, q, qinj, "parity", *wb)
    torch.cuda.synchronize()
    buf = (C.c_longlong * (160 * 16))()
    lib().rn_debug_chain_profile.argtypes = [C.c_void_p]
    assert lib().rn_debug_chain_profile(buf) == 0
    t = torch.tensor(list(buf), dtype=torch.float64).view(160, 16)[:148]
    tiles = t[:, 13].clamp(min=1)
    names = ["issuer wait gen_ready", "issuer wait acc_free", "issuer wait epi_ready[0]", "issuer wait epi_ready[1]",
             "issuer wait w_full", "issuer total", "epi wait acc_full h0 (3/tile)", "epi busy h0 (2/tile)",
             "epi wait acc_full h1 (3/tile)", "epi busy h1 (2/tile)", "epi last layer busy (2 halves)", "gen wait gen_go", "gen busy"]
    print(f"per tile, mean over CTAs (cycles); tiles per CTA {float(tiles.mean()):.1f}; MMA floor 18432")
    for i, nm in enumerate(names):
        print(f"    {nm:34s} {float((t[:, i] / tiles).mean()):10.0f}")


if __name__ == "__main__":
    main()
