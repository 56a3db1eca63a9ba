"""Diagnostics (not a test): CUDA-event timing of rn_relation_fwd / rn_relation_bwd alone at a given batch.

    python tests/diag_tc_time.py [B] [precision] [iters]
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from relationnetworks_clevr_b200 import ops
from tests.test_parity_gpu import _g_params


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    precision = sys.argv[2] if len(sys.argv) > 2 else "parity"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    n, k, Q, G, qinj = 64, 26, 128, 256, 0
    gen = torch.Generator().manual_seed(0)
    x = torch.randn(B, n, k, generator=gen).cuda().requires_grad_(True)
    q = torch.randn(B, Q, generator=gen).cuda().requires_grad_(True)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen, scale=2.0):
        wb += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
    dxg = torch.randn(B, G, generator=gen).cuda()
    ops.timers_enable(True)
    for _ in range(iters):
        xg = ops.RelationFunction.apply(x, q, qinj, precision, *wb)
        xg.backward(dxg)
    t = ops.timers_collect()
    for k_, v in t.items():
        print(f"{k_}: median {statistics.median(v[2:]):.3f} ms  min {min(v[2:]):.3f} ms  (B={B}, {precision})")
    with torch.no_grad():
        ops.timers_enable(True)
        for _ in range(iters):
            ops.RelationFunction.apply(x.detach(), q.detach(), qinj, precision, *[t_.detach() for t_ in wb])
        t = ops.timers_collect()
        v = t["relation_fwd"]
        print(f"relation_fwd(eval): median {statistics.median(v[2:]):.3f} ms  min {min(v[2:]):.3f} ms")


if __name__ == "__main__":
    main()
