"""GPU parity: every C-ABI entry point against the CPU oracle on identical inputs, and the whole
model against the golden fixtures produced by the unmodified reference.

Tolerance metric everywhere: max|got - ref| / max|ref|.
  fp32 (SIMT) kernels: 2e-4 (summation order only)
  parity tcgen05 kernels: 1e-3 (BASELINE.json north_star: "within 1e-3 fp32 relative tolerance")
"""
import pytest
import torch
import torch.nn.functional as F

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from relationnetworks_clevr_b200 import ops
from tests.golden_util import CASES, TRAIN_ONLY_CASES, case_inputs, case_params, load_npz, oracle_train_grads

pytestmark = pytest.mark.gpu

TOL_FP32 = 2e-4
TOL_PARITY = 1e-3
DEV = "cuda"

SHAPES = {
    # name: (B, n, k, Q, G, qinj)
    "fp_d4": (3, 16, 26, 128, 256, 0),
    "fp_d8": (2, 64, 26, 128, 256, 0),
    "ir_d8": (2, 64, 26, 128, 256, 2),
    "sd": (5, 12, 7, 256, 512, 0),
    "ir_sd": (3, 12, 7, 256, 512, 2),
    "odd": (2, 9, 5, 24, 64, 1),
}


def _g_params(n, k, Q, G, qinj, gen, scale=1.0):
    out = []
    for l in range(4):
        fan = (2 * k if l == 0 else G) + (Q if l == qinj else 0)
        w = (torch.rand(G, fan, generator=gen) * 2 - 1) * scale / fan ** 0.5
        b = (torch.rand(G, generator=gen) * 2 - 1) * 0.1
        out.append((w, b))
    return out


def _oracle_grads(x, q, gp, qinj, dxg, dtype):
    def leaf(t):
        return t.detach().clone().to(dtype).requires_grad_(True)

    xx, qq = leaf(x), leaf(q)
    g = [(leaf(w), leaf(b)) for w, b in gp]
    xg = O.g_mlp_dense(xx, qq, g, qinj)
    xg.backward(dxg.to(dtype))
    out = {"xg": xg.detach(), "dx": xx.grad, "dq": qq.grad}
    for l in range(4):
        out[f"dW{l}"], out[f"db{l}"] = g[l][0].grad, g[l][1].grad
    return out


def _relation_case(name, precision, tol):
    """CUDA vs the fp64 oracle.  A ReLU pre-activation that lands within round-off of zero flips its
    gradient mask under ANY fp32 evaluation order (the fp32 CPU oracle shows the same), so each tensor's
    tolerance is floored at 4x the fp32-oracle-vs-fp64-oracle distance measured on the same inputs."""
    B, n, k, Q, G, qinj = SHAPES[name]
    gen = torch.Generator().manual_seed(sorted(SHAPES).index(name) + 11)
    x = torch.randn(B, n, k, generator=gen)
    q = torch.randn(B, Q, generator=gen)
    gp = _g_params(n, k, Q, G, qinj, gen, scale=2.0)
    dxg = torch.randn(B, G, generator=gen)
    ref = _oracle_grads(x, q, gp, qinj, dxg, torch.float64)
    ref32 = _oracle_grads(x, q, gp, qinj, dxg, torch.float32)
    floor = {k_: 4 * O.rel_err(ref32[k_], ref[k_]) for k_ in ref}
    # CUDA
    xc, qc = x.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    wb = []
    for w, b in gp:
        wb += [w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    xg = ops.RelationFunction.apply(xc, qc, qinj, precision, *wb)
    xg.backward(dxg.to(DEV))
    got = {"xg": xg.detach(), "dx": xc.grad, "dq": qc.grad}
    for l in range(4):
        got[f"dW{l}"], got[f"db{l}"] = wb[2 * l].grad, wb[2 * l + 1].grad
    errs = {k_: O.rel_err(got[k_].cpu(), ref[k_]) for k_ in ref}
    print(name, precision, {k_: f"{v:.1e}" for k_, v in errs.items()})
    bad = {k_: (v, floor[k_]) for k_, v in errs.items() if not v <= max(tol, floor[k_])}
    assert not bad, (name, precision, bad)
    return errs


@pytest.mark.parametrize("name", list(SHAPES))
def test_relation_fp32_matches_oracle(name):
    _relation_case(name, "fp32", TOL_FP32)


# tcgen05 path.  Forward (the north-star bar): 1e-3 max-norm vs the fp64 oracle -- measured ~1e-6 .. 1e-5.
# Backward, "parity" mode: the training forward computes every pre-activation as A_hi W_hi + A_lo W_hi + A_hi W_lo (fp16
# splits of activations AND weights, fp32 accumulate), so its ReLU masks agree with an fp32 evaluation; what is left is the
# fp16 rounding of the dZ operands (2^-11 relative, zero mean).  Bars (max-norm, L2-relative), floored like the fp32 tests at
# 4x the fp32-oracle-vs-fp64-oracle distance: dW / db / dq 1e-3; the per-object input gradient dx (a sum of only
# 128 x 256 terms, whose fp32-vs-fp64 distance is itself ~2e-3 on these inputs) 6e-3 / 1.5e-3.
# "fast" mode (one fp16 pass everywhere, fp16 activations): ~1e-4 of the ReLU masks differ from fp32, which shows as
# 2e-4 .. 4e-3 on parameter gradients and 5e-3 .. 2e-2 on dx whatever the batch size (profiles/r02_grad_error_vs_batch.json).
TOL_TC_PARAM = {"parity": (1e-3, 1e-3), "fast": (6e-3, 6e-3)}
TOL_TC_DX = {"parity": (6e-3, 1.5e-3), "fast": (2.5e-2, 1.2e-2)}
TC_CASES = {
    # name: (B, n, qinj)
    "fp_d4_b32": (32, 16, 0),
    "fp_d8_b8": (8, 64, 0),
    "ir_d8_b8": (8, 64, 2),
}


@pytest.mark.parametrize("name", list(TC_CASES))
@pytest.mark.parametrize("precision", ["parity", "fast"])
def test_relation_tcgen05_matches_oracle(name, precision):
    B, n, qinj = TC_CASES[name]
    k, Q, G = 26, 128, 256
    assert ops.tc_supported(n, G, 4, k, Q, qinj)
    gen = torch.Generator().manual_seed(sorted(TC_CASES).index(name) + 31)
    x = torch.randn(B, n, k, generator=gen)
    q = torch.randn(B, Q, generator=gen)
    gp = _g_params(n, k, Q, G, qinj, gen, scale=2.0)
    dxg = torch.randn(B, G, generator=gen)
    ref = _oracle_grads(x, q, gp, qinj, dxg, torch.float64)
    ref32 = _oracle_grads(x, q, gp, qinj, dxg, torch.float32)
    floor = {k_: 4 * O.rel_err(ref32[k_], ref[k_]) for k_ in ref}
    xc, qc = x.to(DEV).requires_grad_(True), q.to(DEV).requires_grad_(True)
    wb = []
    for w, b in gp:
        wb += [w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    xg = ops.RelationFunction.apply(xc, qc, qinj, precision, *wb)
    xg.backward(dxg.to(DEV))
    got = {"xg": xg.detach(), "dx": xc.grad, "dq": qc.grad}
    for l in range(4):
        got[f"dW{l}"], got[f"db{l}"] = wb[2 * l].grad, wb[2 * l + 1].grad
    errs = {}
    for k_ in ref:
        d = got[k_].double().cpu() - ref[k_]
        errs[k_] = (float(d.abs().max() / ref[k_].abs().max()), float(d.norm() / ref[k_].norm()))
    print(name, precision, {k_: f"{a:.1e}/{b_:.1e}" for k_, (a, b_) in errs.items()})
    assert errs["xg"][0] < (TOL_PARITY if precision == "parity" else 3e-3)      # fast: one fp16 pass, ~3e-4 measured
    bad = {}
    for k_, (emax, el2) in errs.items():
        if k_ == "xg":
            continue
        tmax, tl2 = (TOL_TC_DX if k_ == "dx" else TOL_TC_PARAM)[precision]
        if emax > max(tmax, floor[k_]) or el2 > max(tl2, floor[k_]):
            bad[k_] = (emax, el2, floor[k_])
    assert not bad, bad


def test_relation_tcgen05_backward_is_linear_in_dxg():
    """Size-independent property at the FULL bench shape (B=640, d=8): for a fixed forward, backward is linear
    in dxg, so grad(a*d1 + d2) == a*grad(d1) + grad(d2) up to the fp16 operand rounding of the dZ images."""
    B, n, k, Q, G, qinj = 640, 64, 26, 128, 256, 0
    gen = torch.Generator().manual_seed(77)
    x = torch.randn(B, n, k, generator=gen).to(DEV).requires_grad_(True)
    q = torch.randn(B, Q, generator=gen).to(DEV).requires_grad_(True)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen):
        wb += [w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
    d1 = torch.randn(B, G, generator=gen).to(DEV)
    d2 = torch.randn(B, G, generator=gen).to(DEV)
    xg = ops.RelationFunction.apply(x, q, qinj, "parity", *wb)
    leaves = [x, q] + wb
    g1 = torch.autograd.grad(xg, leaves, d1, retain_graph=True)
    g2 = torch.autograd.grad(xg, leaves, d2, retain_graph=True)
    g3 = torch.autograd.grad(xg, leaves, 0.5 * d1 + d2)
    for a, b_, c in zip(g1, g2, g3):
        want = 0.5 * a + b_
        assert float((c - want).abs().max() / want.abs().max()) < 2e-3
    # pair-sum sanity at full size: x_g is finite and matches the fp32 SIMT path
    with torch.no_grad():
        ref = ops.RelationFunction.apply(x, q, qinj, "fp32", *wb)
    assert O.rel_err(xg.detach().cpu(), ref.cpu()) < TOL_PARITY


@pytest.mark.parametrize("stem", ["seeded_original_fp", "ckpt_original_fp"])
def test_gradient_parity_at_bench_shape(stem):
    """The benchmarked configuration (BASELINE.json config 2: original-fp, d = 8, B = 640, train mode, parity precision):
    gradients of all 35 parameter tensors and of the relation op's inputs against this library's fp32 SIMT path, which
    the tests above pin to the CPU oracle at 2e-4.  Bar: 1e-3 max-norm for every parameter gradient and dq (measured
    3e-4 .. 6e-4, profiles/r02_grad_error_vs_batch.json); dx: 1.5e-3 in L2 (max-norm 6e-3: the fp32 conditioning of a
    128 x 256-term sum)."""
    hyp, p = case_params(stem)
    B = 640
    img = O.structured_images(B, 128, 7).to(DEV)
    qst = O.questions(B, 20, 82, 8).to(DEV)
    lab = O.labels(B, 28, 9).to(DEV)
    mask = (torch.rand(B, hyp["f_fc2"], generator=torch.Generator().manual_seed(10)) > 0.5).to(torch.uint8)
    res = {}
    for precision in ("fp32", "parity"):
        m = R.RN(_Args, hyp)
        m.load_state_dict(p, strict=False)
        m.to(DEV).train()
        m.rl.precision = precision
        m.rl.dropout_mask_override = mask
        x = m.conv.objects(img)
        q = m.text(qst)
        x.retain_grad()
        q.retain_grad()
        logp = m.rl(x, q)
        F.nll_loss(logp, lab).backward()
        out = {"logp": logp.detach(), "dx": x.grad, "dq": q.grad}
        out.update({n_: prm.grad for n_, prm in m.named_parameters()})
        res[precision] = {k_: v.double().cpu() for k_, v in out.items()}
        del m
    bad = {}
    for k_, ref in res["fp32"].items():
        if k_.startswith("conv.conv") and k_.endswith("bias"):
            continue
        d = res["parity"][k_] - ref
        emax, el2 = float(d.abs().max() / ref.abs().max()), float(d.norm() / ref.norm())
        tmax, tl2 = (6e-3, 1.5e-3) if k_ == "dx" else (1e-3, 1e-3)
        if emax > tmax or el2 > tl2:
            bad[k_] = (emax, el2)
    assert not bad, bad


def test_relation_eval_forward_and_determinism():
    B, n, k, Q, G, qinj = SHAPES["fp_d8"]
    gen = torch.Generator().manual_seed(5)
    x, q = torch.randn(B, n, k, generator=gen).to(DEV), torch.randn(B, Q, generator=gen).to(DEV)
    wb = []
    for w, b in _g_params(n, k, Q, G, qinj, gen):
        wb += [w.to(DEV), b.to(DEV)]
    for precision in ("fp32", "parity"):
        if precision != "fp32" and not ops.tc_supported(n, G, 4, k, Q, qinj):
            continue
        with torch.no_grad():
            a = ops.RelationFunction.apply(x, q, qinj, precision, *wb)
            b_ = ops.RelationFunction.apply(x, q, qinj, precision, *wb)
        assert torch.equal(a, b_), precision          # no atomics in the pair-sum: bitwise reproducible
        ref = O.g_mlp_dense(x.cpu().double(), q.cpu().double(), [(wb[2 * l].cpu().double(), wb[2 * l + 1].cpu().double()) for l in range(4)], qinj)
        assert O.rel_err(a.cpu(), ref) < (TOL_FP32 if precision == "fp32" else TOL_PARITY)


# 256-wide heads with A <= 32 take the fused one-launch kernels (8 samples per block: 7 and 37 leave a ragged block)
@pytest.mark.parametrize("dims", [(7, 256, 256, 256, 28), (37, 256, 256, 256, 28), (5, 512, 512, 1024, 28)])
@pytest.mark.parametrize("use_mask", [False, True])
def test_f_head_matches_oracle(dims, use_mask):
    B, G, F1, F2, A = dims
    gen = torch.Generator().manual_seed(B + F2)
    p = {"rl.f_fc1.weight": torch.randn(F1, G, generator=gen) / G ** 0.5, "rl.f_fc1.bias": torch.randn(F1, generator=gen) * 0.1,
         "rl.f_fc2.weight": torch.randn(F2, F1, generator=gen) / F1 ** 0.5, "rl.f_fc2.bias": torch.randn(F2, generator=gen) * 0.1,
         "rl.f_fc3.weight": torch.randn(A, F2, generator=gen) / F2 ** 0.5, "rl.f_fc3.bias": torch.randn(A, generator=gen) * 0.1}
    xg = torch.randn(B, G, generator=gen) * 3
    mask = (torch.rand(B, F2, generator=gen) > 0.5) if use_mask else None
    dlogp = torch.randn(B, A, generator=gen)
    p64 = {k_: v.double().requires_grad_(True) for k_, v in p.items()}
    xg64 = xg.double().requires_grad_(True)
    ref = O.f_mlp(p64, xg64, 0.5, use_mask, mask)
    ref.backward(dlogp.double())
    pc = {k_: v.to(DEV).requires_grad_(True) for k_, v in p.items()}
    xgc = xg.to(DEV).requires_grad_(True)
    got = ops.FHeadFunction.apply(xgc, pc["rl.f_fc1.weight"], pc["rl.f_fc1.bias"], pc["rl.f_fc2.weight"], pc["rl.f_fc2.bias"],
                                  pc["rl.f_fc3.weight"], pc["rl.f_fc3.bias"],
                                  mask.to(DEV).to(torch.uint8) if use_mask else None, 2.0)
    got.backward(dlogp.to(DEV))
    assert O.rel_err(got.detach().cpu(), ref.detach()) < TOL_FP32
    assert O.rel_err(xgc.grad.cpu(), xg64.grad) < TOL_FP32
    for k_ in p:
        assert O.rel_err(pc[k_].grad.cpu(), p64[k_].grad) < TOL_FP32, k_


@pytest.mark.parametrize("B,T", [(3, 5), (37, 20), (80, 45), (640, 20)])
def test_question_encoder_matches_oracle(B, T):
    """Embedding + LSTM kernels (csrc/lstm.cu) against the oracle's explicit recurrence in fp64: q and all five gradients."""
    hyp = O.HYPERPARAMS["original-fp"]
    p = {k_: v for k_, v in O.seeded_params(hyp, 82, 28, seed=B + T).items() if k_.startswith("text.")}
    qst = O.questions(B, T, 82, seed=T, left_pad=min(3, T - 1))
    dq = torch.randn(B, 128, generator=torch.Generator().manual_seed(B))
    p64 = {k_: v.double().requires_grad_(True) for k_, v in p.items()}
    ref = O.question_embed(p64, qst)
    ref.backward(dq.double())
    m = R.QuestionEmbedModel(82, embed=32, hidden=128)
    m.load_state_dict({k_[len("text."):]: v for k_, v in p.items()})
    m.to(DEV)
    assert m.use_kernel and ops.lstm_supported(B, T, 83, 32, 128)
    q = m(qst.to(DEV))
    q.backward(dq.to(DEV))
    assert O.rel_err(q.detach().cpu(), ref.detach()) < TOL_FP32
    for name, prm in m.named_parameters():
        assert O.rel_err(prm.grad.cpu(), p64["text." + name].grad) < TOL_FP32, name
    # the PyTorch path (kept for other hidden sizes) agrees too
    m.use_kernel = False
    m.zero_grad()
    q2 = m(qst.to(DEV))
    assert O.rel_err(q2.detach().cpu(), ref.detach()) < TOL_FP32


def _conv_params(seed):
    p = O.seeded_params(O.HYPERPARAMS["original-fp"], 82, 28, seed)
    return {k_: v for k_, v in p.items() if k_.startswith("conv.")}


# side % 64 == 0 runs the tensor-core convolutions (TF32 x3 split: backward by default, flags = 2: forward too), other
# sides the fp32 SIMT kernels; batches that are not multiples of the images-per-block of the 8x8 / 4x4 layers (4 forward,
# 2 weight gradient) exercise the image masks
@pytest.mark.parametrize("B,side", [(3, 32), (2, 64), (5, 64), (2, 128), (5, 128), (1, 192), (1, 256)])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("flags", [0, 2])
def test_conv_objects_match_oracle(B, side, training, flags, monkeypatch):
    if flags and side % 64:
        pytest.skip("the tensor-core kernels need side % 64 == 0")
    monkeypatch.setattr(ops, "conv_flags", flags)
    p = _conv_params(side + B)
    img = O.uniform_images(B, side, seed=side)
    d = side // 16
    dobj = torch.randn(B, d * d, 26, generator=torch.Generator().manual_seed(1))
    p64 = {k_: (v.double().requires_grad_("running" not in k_)) for k_, v in p.items()}
    running, pre = {}, []
    feat = O.conv_features(p64, img.double(), training, running, pre_out=pre)
    obj_ref = O.objects_from_features(feat)
    obj_ref.backward(dobj.double())
    # Conditioning: a ReLU pre-activation within fp32 round-off of zero (|pre| < 1e-7 max|pre|, i.e. ~2 ulp) takes either
    # branch depending on the fp32 evaluation order, and with these tiny batches ONE flipped mask moves the conv1 / conv2
    # weight gradients by 3e-3 .. 3e-2 of their max-norm (sums over few, nearly cancelling terms).  Each such element's
    # effect, measured by re-running the fp64 oracle with that element on the other branch, is added to the tolerance.
    floor = {}
    fragile = [(i + 1, idx) for i, x in enumerate(pre)
               for idx in (x.abs() < 1e-7 * x.abs().max()).nonzero().tolist()]
    assert len(fragile) <= 8, "test inputs are degenerate"
    for layer, idx in fragile:
        mask = torch.zeros_like(pre[layer - 1], dtype=torch.bool)
        mask[tuple(idx)] = True
        q64 = {k_: (v.detach().clone().requires_grad_("running" not in k_)) for k_, v in p64.items()}
        O.objects_from_features(O.conv_features(q64, img.double(), training, flip={layer: mask})).backward(dobj.double())
        for k_, v in q64.items():
            if v.grad is not None and float(p64[k_].grad.abs().max()) > 0:
                floor[k_] = floor.get(k_, 0.0) + float(O.rel_err(v.grad, p64[k_].grad))
    m = R.ConvInputModel()
    m.load_state_dict({k_[len("conv."):]: v for k_, v in p.items()}, strict=False)
    m.to(DEV).train(training)
    obj = m.objects(img.to(DEV))
    obj.backward(dobj.to(DEV))
    assert O.rel_err(obj.detach().cpu(), obj_ref.detach()) < TOL_FP32
    for name, prm in m.named_parameters():
        ref = p64["conv." + name].grad
        if name.startswith("conv") and name.endswith("bias") and training:
            assert float(prm.grad.abs().max()) == 0.0      # exactly zero by construction (BN removes it)
            continue
        assert O.rel_err(prm.grad.cpu(), ref) < 5e-4 + 1.5 * floor.get("conv." + name, 0.0), name
    if training:
        for i in range(1, 5):
            for s in ("running_mean", "running_var"):
                got = getattr(getattr(m, f"batchNorm{i}"), s).cpu()
                assert O.rel_err(got, running[f"conv.batchNorm{i}.{s}"]) < TOL_FP32
    # the reference-shaped [B,24,d,d] view
    with torch.no_grad():
        m.eval()
        f2 = m(img.to(DEV))
        assert f2.shape == (B, 24, d, d)


@pytest.mark.parametrize("B,side", [(13, 128), (3, 192)])
def test_conv_tensor_core_matches_simt(B, side):
    """The tensor-core convolutions (mma.sync TF32 with the 3-pass hi/lo split) against this library's fp32 SIMT kernels on
    the same inputs: both are fp32-level evaluations of the same sums (1e-5 forward, 1e-4 gradients), over enough blocks
    that the persistent kernels loop over several units."""
    p = _conv_params(side + B)
    img = O.uniform_images(B, side, seed=side + 1).to(DEV)
    d = side // 16
    dobj = torch.randn(B, d * d, 26, generator=torch.Generator().manual_seed(3)).to(DEV)
    res = []
    saved_flags = ops.conv_flags
    try:
        for flags in (2, 0, 1):          # tensor-core forward + backward, default (backward only), SIMT
            ops.conv_flags = flags
            m = R.ConvInputModel()
            m.load_state_dict({k_[len("conv."):]: v for k_, v in p.items()}, strict=False)
            m.to(DEV).train()
            obj = m.objects(img)
            obj.backward(dobj)
            res.append({"obj": obj.detach(), **{n: prm.grad for n, prm in m.named_parameters()},
                        **{n: b_.clone() for n, b_ in m.named_buffers() if "running" in n}})
    finally:
        ops.conv_flags = saved_flags
    for r in res[:2]:
        for name in r:
            a, b_ = r[name], res[2][name]
            if float(b_.abs().max()) == 0.0:
                assert float(a.abs().max()) == 0.0, name
                continue
            assert O.rel_err(a.cpu(), b_.cpu()) < (1e-5 if name == "obj" or "running" in name else 1e-4), name
    assert torch.equal(res[1]["obj"], res[2]["obj"])       # the default forward IS the fp32 SIMT forward


@pytest.mark.parametrize("side", [64, 128])
def test_conv_uint8_images_equal_totensor_path(side):
    """Raw uint8 pixels fed straight to the first conv layer (converted u / 255 while staging) give bit-identical objects and
    gradients to the reference's input pipeline, ToTensor() = float().div(255) on the host (train.py:182-188)."""
    B = 3
    p = _conv_params(side)
    img_u8 = torch.randint(0, 256, (B, 3, side, side), dtype=torch.uint8, generator=torch.Generator().manual_seed(side))
    d = side // 16
    dobj = torch.randn(B, d * d, 26, generator=torch.Generator().manual_seed(2)).to(DEV)
    res = []
    for img in (img_u8.float().div(255).to(DEV), img_u8.to(DEV)):
        m = R.ConvInputModel()
        m.load_state_dict({k_[len("conv."):]: v for k_, v in p.items()}, strict=False)
        m.to(DEV).train()
        obj = m.objects(img)
        obj.backward(dobj)
        res.append([obj.detach()] + [prm.grad for prm in m.parameters()])
    for a, b_ in zip(*res):
        assert torch.equal(a, b_)


def test_clip_adam_matches_oracle():
    gen = torch.Generator().manual_seed(9)
    n = 100_003
    w, g = torch.randn(n, generator=gen), torch.randn(n, generator=gen) * 3
    wr, m, v = w.clone(), torch.zeros(n), torch.zeros(n)
    wc, mc, vc = w.to(DEV), torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in range(1, 4):
        total_ref = O.clip_and_adam([wr], [g.clone()], [m], [v], step, lr=1e-3)
        total = ops.clip_adam_(wc, g.to(DEV), mc, vc, step, lr=1e-3)
        assert abs(float(total) - float(total_ref)) < 1e-4 * float(total_ref)
        assert torch.allclose(wc.cpu(), wr, rtol=2e-5, atol=2e-6)


class _Args:
    qdict_size, adict_size = 82, 28


def _build(stem, precision):
    hyp, p = case_params(stem)
    m = R.RN(_Args, hyp)
    m.load_state_dict(p, strict=False)
    m.to(DEV)
    m.rl.precision = precision
    return hyp, m


@pytest.mark.parametrize("stem", list(CASES))
@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_model_eval_matches_reference_golden(stem, precision):
    z = load_npz(stem + "_eval")
    hyp, m = _build(stem, precision)
    img, qst = case_inputs(z)
    m.eval()
    with torch.no_grad():
        logp = m(img.to(DEV), qst.to(DEV)).cpu()
    tol = TOL_FP32 if precision == "fp32" else TOL_PARITY
    assert O.rel_err(logp, torch.from_numpy(z["logp"])) < tol
    assert torch.equal(logp.argmax(1), torch.from_numpy(z["logp"]).argmax(1))


@pytest.mark.parametrize("stem", [s for s in CASES if s != "seeded_original_fp_d12"] + list(TRAIN_ONLY_CASES))
@pytest.mark.parametrize("precision", ["fp32", "auto"])
def test_model_train_step_matches_reference_golden(stem, precision):
    z = load_npz(stem + "_train")
    hyp, m = _build(stem, precision)
    img, qst = case_inputs(z)
    m.train()
    m.rl.dropout_mask_override = torch.from_numpy(z["dropout_mask"]).to(torch.uint8)
    logp = m(img.to(DEV), qst.to(DEV))
    loss = F.nll_loss(logp, torch.from_numpy(z["label"]).to(DEV))
    loss.backward()
    tol = TOL_FP32 if precision == "fp32" else TOL_PARITY
    assert O.rel_err(logp.detach().cpu(), torch.from_numpy(z["logp"])) < tol
    assert abs(float(loss.detach()) - float(z["loss"])) < tol * max(1.0, abs(float(z["loss"])))
    # gradients: against the fp64 oracle (itself pinned to the reference's gradients by
    # tests/test_oracle_golden.py), tolerance floored at 4x the fp32-oracle conditioning noise
    ref, floor = oracle_train_grads(stem)
    bad = {}
    for name, prm in m.named_parameters():
        if name not in ref:
            continue
        if name.startswith("conv.conv") and name.endswith("bias"):
            assert float(prm.grad.abs().max()) == 0.0       # exactly zero under batch statistics
            continue
        err = O.rel_err(prm.grad.cpu(), ref[name])
        # parity mode (3-pass forward: fp32-level ReLU masks) holds the same kind of bar as the fp32 kernels: 1e-3
        # (2e-4 for fp32) floored at 8x the fp32-oracle-vs-fp64-oracle distance of the same tensor on the same inputs.
        # Batch-32 / d = 16 fixtures: 2e-3 for BOTH precisions -- measured worst case 1.2e-3 (fp32 SIMT kernels) and
        # 1.3e-3 (parity) on the trained checkpoint, whose gradients are the least well conditioned (any two fp32
        # evaluation orders disagree on a few ReLU masks; tests/diag_golden_errors.py prints the per-tensor numbers)
        gtol = max(tol if stem in CASES else 2e-3, 8 * floor[name])
        if err > gtol:
            bad[name] = (err, floor[name])
    assert not bad, bad
    for name, buf in m.named_buffers():
        if "running" in name:
            assert O.rel_err(buf.cpu(), torch.from_numpy(z["running/" + name])) < TOL_FP32


@pytest.mark.parametrize("n", [144, 256])
def test_relation_tcgen05_grid_sweep(n):
    """BASELINE.json config 5 (d = 12, 16): tiles straddle `a` boundaries for n = 144; the reference needs 43 GB
    per activation at B=640, d=16, so the check is tcgen05 vs this library's fp32 SIMT path (itself pinned to the
    oracle above) at a small batch."""
    B, k, Q, G, qinj = 2, 26, 128, 256, 0
    assert ops.tc_supported(n, G, 4, k, Q, qinj)
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(B, n, k, generator=gen).to(DEV)
    q = torch.randn(B, Q, generator=gen).to(DEV)
    dxg = torch.randn(B, G, generator=gen).to(DEV)
    params = _g_params(n, k, Q, G, qinj, gen)
    res = {}
    for precision in ("fp32", "parity"):
        xc, qc = x.clone().requires_grad_(True), q.clone().requires_grad_(True)
        wb = []
        for w, b in params:
            wb += [w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)]
        xg = ops.RelationFunction.apply(xc, qc, qinj, precision, *wb)
        xg.backward(dxg)
        res[precision] = [xg.detach(), xc.grad, qc.grad] + [t.grad for t in wb]
    assert O.rel_err(res["parity"][0].cpu(), res["fp32"][0].cpu()) < TOL_PARITY
    assert O.rel_err(res["parity"][1].cpu(), res["fp32"][1].cpu()) < TOL_TC_DX["parity"][0]
    for a, b_ in zip(res["parity"][2:], res["fp32"][2:]):
        assert O.rel_err(a.cpu(), b_.cpu()) < TOL_TC_PARAM["parity"][0]
    # ... and against the CPU oracle (fp64, materialised pairs: 2 x n^2 rows fit on the host at this batch)
    ref = _oracle_grads(x.cpu(), q.cpu(), params, qinj, dxg.cpu(), torch.float64)
    names = ["xg", "dx", "dq"] + [f"d{t}{l}" for l in range(4) for t in ("W", "b")]
    for nm, got in zip(names, res["parity"]):
        tol = TOL_PARITY if nm == "xg" else (TOL_TC_DX if nm == "dx" else TOL_TC_PARAM)["parity"][0]
        assert O.rel_err(got.cpu(), ref[nm]) < tol, (nm, O.rel_err(got.cpu(), ref[nm]))


def test_train_driver_smoke(tmp_path):
    """The reference-surface driver runs an epoch on synthetic CLEVR-shaped batches and writes RN_epoch_01.pth."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "train.py"), "--model", "original-fp", "--epochs", "1", "--batch-size", "32",
           "--synthetic-batches", "3", "--log-interval", "1", "--config", os.path.join(root, "config.json")]
    env = dict(os.environ, PYTHONPATH=root)
    out = subprocess.run(cmd, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Train Epoch: 1 [0/96 (0%)] Train loss:" in out.stdout
    assert "Test Epoch 1: Accuracy" in out.stdout
    ckpts = [os.path.join(dp, f) for dp, _, fs in os.walk(tmp_path) for f in fs if f == "RN_epoch_01.pth"]
    assert len(ckpts) == 1
    sd = torch.load(ckpts[0], map_location="cpu", weights_only=True)
    assert "rl.g_layers.0.weight" in sd and "conv.batchNorm1.running_mean" in sd
