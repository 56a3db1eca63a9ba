"""extraction=True mode and forward hooks (reference extract.py:40-47, 63-86) against fixtures produced by the UNMODIFIED
reference model with its own hooks (tests/golden/make_golden.py: extraction_cases)."""
import contextlib
import io

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from tests.golden_util import case_params, load_npz

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 2e-4


class _Args:
    qdict_size, adict_size = 82, 28


def _build(stem):
    hyp, p = case_params(stem)
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RN(_Args, hyp, extraction=True)
    m.load_state_dict(p, strict=False)
    return hyp, m.to(DEV).eval()


@pytest.mark.parametrize("stem,fixture", [("ckpt_original_fp", "extract_original_fp"), ("ckpt_ir_fp", "extract_ir_fp")])
def test_hooks_and_aggregation_match_reference(stem, fixture):
    z = load_npz(fixture)
    hyp, m = _build(stem)
    B, side, seed = (int(v) for v in z["img_spec"])
    img = O.structured_images(B, side, seed).to(DEV)
    qst = torch.zeros(B, 1, dtype=torch.int64, device=DEV)          # extract.py:102
    qinj, lstm = hyp["question_injection_position"], hyp["lstm_hidden"]
    for idx in range(4):
        got = {}

        def hook(mod, i, o, idx=idx):          # the reference's hook body (extract.py:63-74), on OUR module
            zt = i[0]
            x_ = zt.view(B, zt.size()[0] // B, zt.size()[1])
            if idx == qinj:
                x_ = x_[:, :, :zt.size()[1] - lstm]
            x_ = F.normalize(x_, p=2, dim=2)
            got["max"], got["avg"] = x_.max(1)[0].squeeze(), x_.mean(1).squeeze()

        h = m.rl.g_layers[idx].register_forward_hook(hook)
        with torch.no_grad():
            assert m(img, qst) is None                 # extraction mode stops after g, like the reference
        h.remove()
        assert O.rel_err(got["max"].cpu(), torch.from_numpy(z[f"g{idx}/max"])) < TOL, idx
        assert O.rel_err(got["avg"].cpu(), torch.from_numpy(z[f"g{idx}/avg"])) < TOL, idx
        # the fused aggregation kernel (no normalised copy of the [B*n*n, W] tensor)
        with torch.no_grad():
            x = m.conv.objects(img)
            q = m.text(qst)
            maxf, avgf = m.rl.extract(x, q, idx)
        assert O.rel_err(maxf.cpu(), torch.from_numpy(z[f"g{idx}/max"])) < TOL, idx
        assert O.rel_err(avgf.cpu(), torch.from_numpy(z[f"g{idx}/avg"])) < TOL, idx
    got = {}
    h = m.conv.register_forward_hook(lambda mod, i, o: got.__setitem__("o", o.detach().clone()))      # extract.py:75-86
    with torch.no_grad():
        m(img, qst)
    h.remove()
    o = got["o"].reshape(B, 24, 64)
    assert O.rel_err(o.mean(2).cpu(), torch.from_numpy(z["conv/avg"])) < TOL
    assert O.rel_err(o.max(2)[0].cpu(), torch.from_numpy(z["conv/max"])) < TOL


def test_unhooked_model_is_unaffected():
    hyp, p = case_params("ckpt_original_fp")
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RN(_Args, hyp)
    m.load_state_dict(p, strict=False)
    m.to(DEV).eval()
    img = O.structured_images(2, 128, 3).to(DEV)
    qst = O.questions(2, 20, 82, 4).to(DEV)
    with torch.no_grad():
        a = m(img, qst)
        h = m.rl.g_layers[2].register_forward_hook(lambda mod, i, o: None)
        b = m(img, qst)          # hooked: materialised fp32 path
        h.remove()
        c = m(img, qst)
    assert torch.equal(a, c)
    assert O.rel_err(b.cpu(), a.cpu()) < 1e-3
