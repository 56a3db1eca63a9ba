"""Diagnostics (not a test): conv gradients of a golden training case with the tensor-core and the SIMT convolutions."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from relationnetworks_clevr_b200 import ops
from tests.golden_util import case_inputs, case_params, load_npz, oracle_train_grads
from tests.test_parity_gpu import _build


def run(stem, precision, flags):
    ops.conv_flags = flags
    z = load_npz(stem + "_train")
    hyp, m = _build(stem, precision)
    img, qst = case_inputs(z)
    m.train()
    m.rl.dropout_mask_override = torch.from_numpy(z["dropout_mask"]).to(torch.uint8)
    logp = m(img.cuda(), qst.cuda())
    F.nll_loss(logp, torch.from_numpy(z["label"]).cuda()).backward()
    torch.cuda.synchronize()
    return {n: p.grad.detach().cpu().clone() for n, p in m.named_parameters() if n.startswith("conv.")}


for stem in sys.argv[1:] or ["seeded_original_fp", "seeded_original_fp_b32"]:
    ref, floor = oracle_train_grads(stem)
    for precision in ("fp32", "auto"):
        for rep in range(2):
            a, b = run(stem, precision, 0), run(stem, precision, 1)
            print(stem, precision, rep, " ".join(
                f"{n.split('.')[1]}:{O.rel_err(a[n], ref[n]):.1e}/{O.rel_err(b[n], ref[n]):.1e}" for n in a if n.endswith("weight") and "conv.conv" in n))
