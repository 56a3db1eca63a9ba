"""Diagnostics (not a test): gradient error of the tcgen05 precision modes against this library's fp32 SIMT path
(pinned to the CPU oracle at 2e-4 by tests/test_parity_gpu.py) as a function of the batch size, whole model,
original-fp, d = 8, train mode with a fixed dropout mask.  Writes one JSON record per (weights, batch, mode).

    python tests/diag_grad_curve.py [out.json] [modes, comma separated] [batches, comma separated]

Metric: max|got - ref| / max|ref| per tensor, and the relative L2 distance.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import contextlib
import io

import torch
import torch.nn.functional as F

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from tests.golden_util import case_params

DEV = "cuda"


class _Args:
    qdict_size, adict_size = 82, 28


def run(stem, B, precision, seed=7):
    hyp, p = case_params(stem)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RN(_Args, hyp)
    m.load_state_dict(p, strict=False)
    m.to(DEV).train()
    m.rl.precision = precision
    img = O.structured_images(B, 128, seed).to(DEV)
    qst = O.questions(B, 20, 82, seed + 1).to(DEV)
    lab = O.labels(B, 28, seed + 2).to(DEV)
    mask = (torch.rand(B, hyp["f_fc2"], generator=torch.Generator().manual_seed(seed + 3)) > 0.5).to(torch.uint8)
    m.rl.dropout_mask_override = mask
    x = m.conv.objects(img)
    q = m.text(qst)
    x.retain_grad()
    q.retain_grad()
    logp = m.rl(x, q)
    loss = F.nll_loss(logp, lab)
    loss.backward()
    torch.cuda.synchronize()
    out = {"logp": logp.detach().double().cpu(), "dx": x.grad.double().cpu(), "dq": q.grad.double().cpu()}
    for name, prm in m.named_parameters():
        out[name] = prm.grad.double().cpu()
    return out


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/grad_curve.json"
    modes = (sys.argv[2] if len(sys.argv) > 2 else "parity").split(",")
    batches = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "8,32,160,640").split(",")]
    records = []
    for stem in ("seeded_original_fp", "ckpt_original_fp"):
        for B in batches:
            ref = run(stem, B, "fp32")
            for mode in modes:
                got = run(stem, B, mode)
                errs = {}
                for k, r in ref.items():
                    if k.startswith("conv.conv") and k.endswith("bias"):
                        continue                     # exactly zero under batch statistics
                    d = got[k] - r
                    errs[k] = (float(d.abs().max() / r.abs().max().clamp_min(1e-300)), float(d.norm() / r.norm().clamp_min(1e-300)))
                groups = {"logp": ["logp"], "dx": ["dx"], "dq": ["dq"],
                          "g_dW": [k for k in errs if k.startswith("rl.g_layers") and k.endswith("weight")],
                          "g_db": [k for k in errs if k.startswith("rl.g_layers") and k.endswith("bias")],
                          "f": [k for k in errs if k.startswith("rl.f_")],
                          "conv": [k for k in errs if k.startswith("conv.")],
                          "text": [k for k in errs if k.startswith("text.")]}
                summary = {g: [max(errs[k][0] for k in ks), max(errs[k][1] for k in ks)] for g, ks in groups.items() if ks}
                rec = {"weights": stem, "B": B, "mode": mode, "env_dgrad_passes": os.environ.get("RN_B200_DGRAD_PASSES", "1"),
                       "max_norm_and_l2": summary, "per_tensor": errs}
                records.append(rec)
                print(stem, B, mode, {g: f"{a:.1e}/{b:.1e}" for g, (a, b) in summary.items()}, flush=True)
    with open(out_path, "w") as f:
        json.dump(records, f, indent=1)


if __name__ == "__main__":
    main()
