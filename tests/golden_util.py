"""Load the golden fixtures written by tests/golden/make_golden.py (reference outputs)."""
import os

import numpy as np
import torch

from oracle import rn_oracle as O

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
QDICT, ADICT = 82, 28

# fixture stem -> (config, numpy seed of the parameters, or the *_eval fixture that stores them)
CASES = {
    "ckpt_original_fp": ("original-fp", "ckpt_original_fp_eval"),
    "ckpt_ir_fp": ("ir-fp", "ckpt_ir_fp_eval"),
    "seeded_original_fp": ("original-fp", 101),
    "seeded_ir_fp": ("ir-fp", 102),
    "seeded_original_sd": ("original-sd", 201),
    "seeded_ir_sd": ("ir-sd", 202),
    "seeded_original_fp_d12": ("original-fp", 301),
}


# round-2 training fixtures (no *_eval twin): batch 32 and the d = 16 grid
TRAIN_ONLY_CASES = {
    "ckpt_original_fp_b32": ("original-fp", "ckpt_original_fp_eval"),
    "seeded_original_fp_b32": ("original-fp", 401),
    "seeded_ir_fp_b32": ("ir-fp", 411),
    "seeded_original_fp_d16": ("original-fp", 421),
}


def load_npz(name):
    return np.load(os.path.join(GOLDEN_DIR, name + ".npz"))


def case_params(stem, dtype=torch.float32):
    config, src = CASES[stem] if stem in CASES else TRAIN_ONLY_CASES[stem]
    hyp = O.HYPERPARAMS[config]
    if isinstance(src, int):
        return hyp, O.seeded_params(hyp, QDICT, ADICT, src, dtype)
    z = load_npz(src)
    return hyp, {k[len("param/"):]: torch.from_numpy(z[k]).to(dtype) for k in z.files if k.startswith("param/")}


def case_inputs(z, dtype=torch.float32):
    kind = str(z["img_kind"])
    B, side, seed = (int(v) for v in z["img_spec"])
    if kind == "structured":
        img = O.structured_images(B, side, seed)
    elif kind == "uniform":
        img = O.uniform_images(B, side, seed)
    else:
        img = O.state_descriptions(B, seed)
    assert abs(float(img.double().sum()) - float(z["img_sum"])) < 1e-6 * max(1.0, abs(float(z["img_sum"])))
    return img.to(dtype), torch.from_numpy(z["qst"])


def golden_grad_check(z, name, grad, tol, errs):
    """Compare one gradient tensor with the (possibly sampled) golden record.

    conv biases feed a train-mode BatchNorm, so their true gradient is exactly zero and the
    reference's value is round-off noise: they are compared on the scale of the same layer's
    weight gradient instead of their own."""
    floor = 0.0
    if name.startswith("conv.conv") and name.endswith(".bias"):
        floor = float(z[f"grad/{name[:-4]}weight/l2"])
    g = grad.detach().float().contiguous().view(-1).cpu()
    key = f"grad/{name}/"
    if key + "full" in z.files:
        ref = torch.from_numpy(z[key + "full"])
        got = g
    else:
        stride = int(z[key + "stride"])
        ref = torch.from_numpy(z[key + "sample"])
        got = g[::stride][: ref.numel()]
    scale = float(z[key + "l2"]) / max(1.0, g.numel() ** 0.5)      # rms of the full tensor
    denom = max(float(ref.abs().max()), scale, floor, 1e-30)
    err = float((got.double() - ref.double()).abs().max()) / denom
    errs[name] = err
    assert err <= tol, (name, err)
    l2 = float(g.double().norm())
    assert abs(l2 - float(z[key + "l2"])) <= tol * max(float(z[key + "l2"]), floor, 1e-30) * 4, \
        (name, l2, float(z[key + "l2"]))


_grad_cache = {}


def oracle_train_grads(stem):
    """fp64 oracle gradients of a *_train fixture plus, per tensor, the conditioning floor: the distance
    between the fp32 and fp64 oracle on the same inputs (ReLU-mask flips at |z| ~ round-off make some
    cases ill-conditioned for ANY fp32 evaluation order).  Cached per process."""
    if stem in _grad_cache:
        return _grad_cache[stem]
    import torch.nn.functional as F

    z = load_npz(stem + "_train")
    hyp, p = case_params(stem)
    img, qst = case_inputs(z)
    lab = torch.from_numpy(z["label"])
    mask = torch.from_numpy(z["dropout_mask"])
    grads = {}
    for dt in (torch.float64, torch.float32):
        leaves = {k: v.to(dt).clone().requires_grad_("running" not in k) for k, v in p.items()}
        F.nll_loss(O.rn_forward(leaves, hyp, img.to(dt), qst, True, mask), lab).backward()
        grads[dt] = {k: v.grad for k, v in leaves.items() if v.grad is not None}
    floor = {k: O.rel_err(grads[torch.float32][k], g) for k, g in grads[torch.float64].items()}
    _grad_cache[stem] = (grads[torch.float64], floor)
    return _grad_cache[stem]
