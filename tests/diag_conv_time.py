"""Diagnostics (not a test): conv stack forward + backward alone at batch B (CUDA-event timing)."""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import relationnetworks_clevr_b200 as R
from relationnetworks_clevr_b200 import ops


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    torch.manual_seed(0)
    conv = R.ConvInputModel().cuda().train()
    img = torch.rand(B, 3, 128, 128, device="cuda")
    dobj = torch.randn(B, 64, 26, device="cuda")
    ops.timers_enable(True)
    for _ in range(iters):
        obj = conv.objects(img)
        obj.backward(dobj)
        for p in conv.parameters():
            p.grad = None
    t = ops.timers_collect()
    for k_, v in t.items():
        print(f"{k_}: median {statistics.median(v[2:]):.3f} ms  min {min(v[2:]):.3f} ms  (B={B})")


if __name__ == "__main__":
    main()
