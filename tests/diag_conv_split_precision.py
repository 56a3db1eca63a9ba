"""Diagnostics (CPU, not a test): would a tensor-core conv stack meet the fp32 parity bar?

Emulates fp16-split operands with fp32 accumulation (what a tcgen05 implicit-GEMM conv would compute) for the
4 x [conv3x3 s2 -> BatchNorm(batch stats) -> ReLU] extractor, against the fp64 oracle:
  1 pass : x_hi * w_hi          2 pass : + x_hi*w_lo (weights split) or + x_lo*w_hi (activations split)
  3 pass : x_hi*w_hi + x_hi*w_lo + x_lo*w_hi
Metric: max|got - ref| / max|ref| on the [B, 24, 8, 8] feature map (tests use 2e-4 for the fp32 SIMT kernels).
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import rn_oracle as O


def split(t):
    hi = t.half().float()
    lo = (t - hi).half().float()
    return hi, lo


def conv_stack(p, img, passes, dtype):
    x = img.to(dtype)
    for l in range(1, 5):
        w, b = p[f"conv.conv{l}.weight"].to(dtype), p[f"conv.conv{l}.bias"].to(dtype)
        if passes == 0:
            y = F.conv2d(x, w, b, stride=2, padding=1)
        else:
            xh, xl = split(x.float())
            wh, wl = split(w.float())
            y = F.conv2d(xh, wh, None, stride=2, padding=1)
            if passes == 3:
                y = y + F.conv2d(xh, wl, None, stride=2, padding=1) + F.conv2d(xl, wh, None, stride=2, padding=1)
            elif passes == 2:          # weight split only
                y = y + F.conv2d(xh, wl, None, stride=2, padding=1)
            elif passes == -2:         # activation split only
                y = y + F.conv2d(xl, wh, None, stride=2, padding=1)
            y = (y + b.float().view(1, -1, 1, 1)).to(dtype)
        g, be = p[f"conv.batchNorm{l}.weight"].to(dtype), p[f"conv.batchNorm{l}.bias"].to(dtype)
        mean = y.mean(dim=(0, 2, 3), keepdim=True)
        var = y.var(dim=(0, 2, 3), unbiased=False, keepdim=True)
        x = torch.relu((y - mean) / torch.sqrt(var + 1e-5) * g.view(1, -1, 1, 1) + be.view(1, -1, 1, 1))
    return x


if __name__ == "__main__":
    hyp = O.HYPERPARAMS["original-fp"]
    for name, img in (("uniform", O.uniform_images(16, 128, 1)), ("structured", O.structured_images(16, 128, 2))):
        for seed in (5, 6):
            p = O.seeded_params(hyp, 82, 28, seed=seed)
            ref = conv_stack(p, img, 0, torch.float64)
            line = {}
            for tag, passes, dt in (("fp32", 0, torch.float32), ("1-pass fp16", 1, torch.float32), ("2-pass w split", 2, torch.float32),
                                    ("2-pass x split", -2, torch.float32), ("3-pass fp16 split", 3, torch.float32)):
                line[tag] = f"{O.rel_err(conv_stack(p, img, passes, dt).double(), ref):.1e}"
            print(name, "seed", seed, line)
