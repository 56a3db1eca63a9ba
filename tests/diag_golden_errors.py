import sys, torch, torch.nn.functional as F
sys.path.insert(0,'/root/repo')
import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from tests.golden_util import case_inputs, case_params, load_npz, oracle_train_grads
class A: qdict_size, adict_size = 82, 28
for stem in ("ckpt_original_fp_b32","seeded_original_fp_b32","seeded_original_fp_d16","ckpt_original_fp"):
    z = load_npz(stem + "_train")
    ref, floor = oracle_train_grads(stem)
    for precision in ("fp32","parity"):
        hyp, p = case_params(stem)
        m = R.RN(A, hyp); m.load_state_dict(p, strict=False); m.cuda().train(); m.rl.precision = precision
        img, qst = case_inputs(z)
        m.rl.dropout_mask_override = torch.from_numpy(z["dropout_mask"]).to(torch.uint8)
        logp = m(img.cuda(), qst.cuda())
        F.nll_loss(logp, torch.from_numpy(z["label"]).cuda()).backward()
        errs = {n: O.rel_err(prm.grad.cpu(), ref[n]) for n, prm in m.named_parameters() if n in ref and not (n.startswith("conv.conv") and n.endswith("bias"))}
        worst = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
        print(stem, precision, "logp", f"{O.rel_err(logp.detach().cpu(), torch.from_numpy(z['logp'])):.1e}", [(n, f"{e:.1e}", f"floor {floor[n]:.1e}") for n, e in worst], flush=True)
