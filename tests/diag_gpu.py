"""Diagnostics (not a test): per-stage error of the CUDA model against the golden fixtures / fp64 oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from tests.golden_util import CASES, case_inputs, case_params, load_npz

class A: qdict_size, adict_size = 82, 28

def stage_errs(stem, precision, train):
    z = load_npz(stem + ("_train" if train else "_eval"))
    hyp, p = case_params(stem)
    img, qst = case_inputs(z)
    m = R.RN(A, hyp); m.load_state_dict(p, strict=False); m.cuda().train(train); m.rl.precision = precision
    p64 = {k: v.double().requires_grad_(train and "running" not in k) for k, v in p.items()}
    mask = torch.from_numpy(z["dropout_mask"]) if train else None
    ref, parts = O.rn_forward(p64, hyp, img.double(), qst, train, mask, return_parts=True)
    out = {}
    with torch.set_grad_enabled(train):
        x = img.cuda() if hyp["state_description"] else m.conv.objects(img.cuda())
        q = m.text(qst.cuda())
        xg = m.rl.relation(x, q)
        if train: m.rl.dropout_mask_override = mask.to(torch.uint8)
        logp = m.rl(x, q)
    out["x"] = O.rel_err(x.detach().cpu(), parts["x"].detach())
    out["q"] = O.rel_err(q.detach().cpu(), parts["q"].detach())
    out["xg"] = O.rel_err(xg.detach().cpu(), parts["x_g"].detach())
    out["logp"] = O.rel_err(logp.detach().cpu(), ref.detach())
    out["logp_vs_golden"] = O.rel_err(logp.detach().cpu(), torch.from_numpy(z["logp"]))
    out["golden_vs_fp64"] = O.rel_err(torch.from_numpy(z["logp"]), ref.detach())
    if train:
        lab = torch.from_numpy(z["label"])
        F.nll_loss(logp, lab.cuda()).backward()
        F.nll_loss(ref, lab).backward()
        worst = {}
        for name, prm in m.named_parameters():
            if prm.grad is None or p64[name].grad is None: continue
            g, r = prm.grad.double().cpu(), p64[name].grad
            worst[name] = (O.rel_err(g, r), float((g - r).norm() / (r.norm() + 1e-300)))
        top = sorted(worst.items(), key=lambda kv: -kv[1][0])[:6]
        out["grads_worst(max,l2)"] = {k: (f"{a:.1e}", f"{b:.1e}") for k, (a, b) in top}
    return out

if __name__ == "__main__":
    precision = sys.argv[1] if len(sys.argv) > 1 else "fp32"
    for stem in CASES:
        for train in (False, True):
            if train and stem.endswith("d12"): continue
            try:
                e = stage_errs(stem, precision, train)
                print(stem, "train" if train else "eval", {k: (f"{v:.1e}" if isinstance(v, float) else v) for k, v in e.items()}, flush=True)
            except Exception as ex:
                print(stem, train, "ERROR", repr(ex)[:300], flush=True)
