import os, sys, contextlib, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import relationnetworks_clevr_b200 as R
from relationnetworks_clevr_b200 import ops
from oracle import rn_oracle as O
from tests.golden_util import case_inputs, case_params, load_npz
class A: qdict_size, adict_size = 82, 28
DEV = "cuda"
def run(stem, aux, precision):
    ops.use_aux_stream = aux
    z = load_npz(stem + "_train")
    hyp, p = case_params(stem)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RN(A, hyp)
    m.load_state_dict(p, strict=False); m.to(DEV); m.rl.precision = precision
    img, qst = case_inputs(z)
    m.train()
    m.rl.dropout_mask_override = torch.from_numpy(z["dropout_mask"]).to(torch.uint8)
    logp = m(img.to(DEV), qst.to(DEV))
    loss = F.nll_loss(logp, torch.from_numpy(z["label"]).to(DEV))
    loss.backward()
    return {n: prm.grad.cpu().clone() for n, prm in m.named_parameters()}
for stem in ("ckpt_original_fp", "ckpt_ir_fp", "seeded_original_fp"):
    for precision in ("fp32", "auto"):
        ref = run(stem, False, precision)
        for it in range(2):
            got = run(stem, True, precision)
            bad = {n: float((got[n] - ref[n]).abs().max() / ref[n].abs().max().clamp_min(1e-30)) for n in ref}
            worst = sorted(bad.items(), key=lambda kv: -kv[1])[:3]
            print(stem, precision, it, [(n, f"{e:.2e}") for n, e in worst], flush=True)
