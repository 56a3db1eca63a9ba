"""Diagnostics (not a test): BASELINE.json config 5 -- relation op forward+backward over the grid sweep
n in {64, 144, 256} (d = 8, 12, 16) at constant pair-row counts, with the achieved algorithmic TFLOP/s."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from relationnetworks_clevr_b200 import ops
from tests.test_parity_gpu import _g_params

FLOP_PER_PAIR_TRAIN = 3 * 2 * 242_688


def run(B, n, precision="parity", iters=6):
    k, Q, G, qinj = 26, 128, 256, 0
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(B, n, k, generator=gen).cuda().requires_grad_(True)
    q = torch.randn(B, Q, generator=gen).cuda().requires_grad_(True)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen, scale=2.0):
        wb += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
    dxg = torch.randn(B, G, generator=gen).cuda()
    ops.timers_enable(True)
    for _ in range(iters):
        ops.RelationFunction.apply(x, q, qinj, precision, *wb).backward(dxg)
    t = ops.timers_collect()
    f, b_ = statistics.median(t["relation_fwd"][2:]), statistics.median(t["relation_bwd"][2:])
    tf = FLOP_PER_PAIR_TRAIN * B * n * n / ((f + b_) * 1e-3) / 1e12
    print(f"n={n:4d} (d={int(n ** 0.5):2d}) B={B:4d} pairs/sample={n * n:6d}: fwd {f:.3f} ms  bwd {b_:.3f} ms  -> {tf:.0f} TFLOP/s algorithmic")


if __name__ == "__main__":
    for B, n in ((640, 64), (128, 144), (40, 256)):
        run(B, n)
