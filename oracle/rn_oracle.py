"""CPU oracle for the Relation-Network hot path.  TEST INFRASTRUCTURE ONLY.

This module is a plain-PyTorch, CPU, functional restatement of the algorithm in the
reference's ``model.py`` (mesnico/RelationNetworks-CLEVR).  It exists to *check* the CUDA
path; it is never imported by the product package (``relationnetworks_clevr_b200``).  Only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import it.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the unmodified reference
``model.py`` from ``/root/reference`` (it runs under torch 2.11 on CPU), drives it with
seeded inputs and both shipped checkpoints, and stores inputs/outputs/gradients in
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function here against
those vectors.

All arithmetic is whatever dtype the inputs carry (fp32 for parity with the reference,
fp64 for a high-precision yardstick).  Parameters are passed as a flat ``dict`` keyed by
the reference's ``state_dict`` names (``conv.conv1.weight`` ... ``rl.g_layers.3.bias``).

Two formulations of the relation layer are provided:

* ``relation_pairs`` / ``g_mlp_dense``: the reference's literal algorithm -- materialise all
  n*n ordered pairs, 4 dense layers (model.py:104-152).  This is what the CPU baseline times.
* ``g_mlp_factorised``: the algebraically identical form the CUDA kernels use (layer 0 split
  into per-object projections U, V and a per-sample bias), with
  hand-written backward formulas (``g_backward_factorised``) so every intermediate the
  kernels produce has an oracle value.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]


# --------------------------------------------------------------------------------------
# hyper-parameters (config.json of the reference; reference config.json:3-58)
# --------------------------------------------------------------------------------------
HYPERPARAMS = {
    "original-fp": dict(state_description=False, g_layers=[256, 256, 256, 256], question_injection_position=0,
                        f_fc1=256, f_fc2=256, dropout=0.5, lstm_hidden=128, lstm_word_emb=32, rl_in_size=52),
    "original-sd": dict(state_description=True, g_layers=[512, 512, 512, 512], question_injection_position=0,
                        f_fc1=512, f_fc2=1024, dropout=0.05, lstm_hidden=256, lstm_word_emb=32, rl_in_size=14),
    "ir-fp": dict(state_description=False, g_layers=[256, 256, 256, 256], question_injection_position=2,
                  f_fc1=256, f_fc2=256, dropout=0.5, lstm_hidden=128, lstm_word_emb=32, rl_in_size=52),
    "ir-sd": dict(state_description=True, g_layers=[512, 512, 512, 512], question_injection_position=2,
                  f_fc1=512, f_fc2=1024, dropout=0.05, lstm_hidden=256, lstm_word_emb=32, rl_in_size=14),
}


def param_shapes(hyp: dict, qdict_size: int, adict_size: int) -> Dict[str, Tuple[int, ...]]:
    """Shapes of every learnable tensor, in the reference's registration order
    (model.py:13-20 conv, :43-44 text, :64-66 f, :88-101 g)."""
    shapes: Dict[str, Tuple[int, ...]] = {}
    cin = 3
    for i in range(1, 5):
        shapes[f"conv.conv{i}.weight"] = (24, cin, 3, 3)
        shapes[f"conv.conv{i}.bias"] = (24,)
        shapes[f"conv.batchNorm{i}.weight"] = (24,)
        shapes[f"conv.batchNorm{i}.bias"] = (24,)
        cin = 24
    E, H = hyp["lstm_word_emb"], hyp["lstm_hidden"]
    shapes["text.wembedding.weight"] = (qdict_size + 1, E)
    shapes["text.lstm.weight_ih_l0"] = (4 * H, E)
    shapes["text.lstm.weight_hh_l0"] = (4 * H, H)
    shapes["text.lstm.bias_ih_l0"] = (4 * H,)
    shapes["text.lstm.bias_hh_l0"] = (4 * H,)
    G = hyp["g_layers"]
    shapes["rl.f_fc1.weight"] = (hyp["f_fc1"], G[-1])
    shapes["rl.f_fc1.bias"] = (hyp["f_fc1"],)
    shapes["rl.f_fc2.weight"] = (hyp["f_fc2"], hyp["f_fc1"])
    shapes["rl.f_fc2.bias"] = (hyp["f_fc2"],)
    shapes["rl.f_fc3.weight"] = (adict_size, hyp["f_fc2"])
    shapes["rl.f_fc3.bias"] = (adict_size,)
    for l, width in enumerate(G):
        fan_in = hyp["rl_in_size"] if l == 0 else G[l - 1]
        if l == hyp["question_injection_position"]:
            fan_in += H
        shapes[f"rl.g_layers.{l}.weight"] = (width, fan_in)
        shapes[f"rl.g_layers.{l}.bias"] = (width,)
    return shapes


def buffer_shapes() -> Dict[str, Tuple[int, ...]]:
    out = {}
    for i in range(1, 5):
        out[f"conv.batchNorm{i}.running_mean"] = (24,)
        out[f"conv.batchNorm{i}.running_var"] = (24,)
    return out


def seeded_params(hyp: dict, qdict_size: int, adict_size: int, seed: int, dtype=torch.float32) -> Params:
    """Deterministic parameters from a numpy PCG64 stream (stable across library versions), so
    golden fixtures need not store weights.  Uniform(-1/sqrt(fan_in), 1/sqrt(fan_in)) like
    torch's default Linear/Conv init scale; BN gamma in [0.5, 1.5], beta in [-0.5, 0.5]; running
    stats non-trivial so eval mode exercises them."""
    import numpy as np

    rng = np.random.default_rng(seed)
    p: Params = {}
    for name, shp in param_shapes(hyp, qdict_size, adict_size).items():
        if "batchNorm" in name:
            lo, hi = (0.5, 1.5) if name.endswith("weight") else (-0.5, 0.5)
        elif name == "text.wembedding.weight":
            lo, hi = -1.0, 1.0
        else:
            if name.startswith("text.lstm"):
                fan_in = hyp["lstm_hidden"]
            elif len(shp) == 1:
                # bias: use the fan-in of the matching weight
                wname = name[: -len("bias")] + "weight"
                wshape = param_shapes(hyp, qdict_size, adict_size)[wname]
                fan_in = int(np.prod(wshape[1:]))
            else:
                fan_in = int(np.prod(shp[1:]))
            b = 1.0 / math.sqrt(fan_in)
            lo, hi = -b, b
        p[name] = torch.from_numpy(rng.uniform(lo, hi, size=shp)).to(dtype)
    for name, shp in buffer_shapes().items():
        if name.endswith("running_mean"):
            p[name] = torch.from_numpy(rng.uniform(-0.3, 0.3, size=shp)).to(dtype)
        else:
            p[name] = torch.from_numpy(rng.uniform(0.5, 1.5, size=shp)).to(dtype)
    return p


# --------------------------------------------------------------------------------------
# conv feature extractor (reference model.py:9-36)
# --------------------------------------------------------------------------------------
BN_EPS = 1e-5       # nn.BatchNorm2d default, model.py:14
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


def conv_features(p: Params, img: torch.Tensor, training: bool,
                  running_out: Optional[Params] = None, pre_out: Optional[list] = None,
                  flip: Optional[dict] = None) -> torch.Tensor:
    """4 x [conv3x3 stride 2 pad 1 -> BatchNorm(24) -> ReLU]  (model.py:22-36).

    training=True normalises with biased batch statistics and, if ``running_out`` is given,
    stores the updated running stats there (momentum 0.1, unbiased variance) as
    nn.BatchNorm2d does.  [B,3,S,S] -> [B,24,S/16,S/16].

    Conditioning diagnostics for the parity tests (not part of the reference's arithmetic): ``pre_out`` collects
    the four ReLU pre-activations; ``flip`` = {layer (1-based): bool mask} evaluates the listed elements with the
    OTHER branch of the ReLU (value min(x, 0) ~ 0, derivative 1 - 1[x > 0]) -- what any evaluation whose round-off
    moves a pre-activation across zero computes."""
    x = img
    for i in range(1, 5):
        x = F.conv2d(x, p[f"conv.conv{i}.weight"], p[f"conv.conv{i}.bias"], stride=2, padding=1)
        gamma, beta = p[f"conv.batchNorm{i}.weight"], p[f"conv.batchNorm{i}.bias"]
        if training:
            mean = x.mean(dim=(0, 2, 3))
            var = x.var(dim=(0, 2, 3), unbiased=False)
            if running_out is not None:
                cnt = x.numel() // x.shape[1]
                rm, rv = p[f"conv.batchNorm{i}.running_mean"], p[f"conv.batchNorm{i}.running_var"]
                running_out[f"conv.batchNorm{i}.running_mean"] = ((1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean).detach()
                running_out[f"conv.batchNorm{i}.running_var"] = (
                    (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var * (cnt / max(cnt - 1, 1))).detach()
        else:
            mean = p[f"conv.batchNorm{i}.running_mean"]
            var = p[f"conv.batchNorm{i}.running_var"]
        x = (x - mean[None, :, None, None]) * torch.rsqrt(var[None, :, None, None] + BN_EPS)
        x = x * gamma[None, :, None, None] + beta[None, :, None, None]
        if pre_out is not None:
            pre_out.append(x.detach())
        if flip is not None and i in flip:
            x = torch.where(flip[i], x - torch.relu(x), torch.relu(x))
        else:
            x = torch.relu(x)
    return x


def coord_grid(d: int, dtype=torch.float32) -> torch.Tensor:
    """[2, d*d] coordinate channels (model.py:208-213): linspace(-d/2, d/2, d); channel 0
    varies along the last (W) axis, channel 1 along H; cell index = row*d + col."""
    c = torch.linspace(-d / 2.0, d / 2.0, d, dtype=dtype)
    xs = c.unsqueeze(0).expand(d, d)
    ys = c.unsqueeze(1).expand(d, d)
    return torch.stack((xs, ys)).reshape(2, d * d)


def objects_from_features(feat: torch.Tensor) -> torch.Tensor:
    """[B,24,d,d] conv output -> [B, d*d, 26] objects with coords appended (model.py:192-201)."""
    b, k, d, _ = feat.shape
    x = feat.reshape(b, k, d * d)
    coords = coord_grid(d, feat.dtype).unsqueeze(0).expand(b, 2, d * d)
    return torch.cat([x, coords], dim=1).permute(0, 2, 1)


# --------------------------------------------------------------------------------------
# question encoder (reference model.py:39-58)
# --------------------------------------------------------------------------------------
def question_embed(p: Params, qst_idxs: torch.Tensor) -> torch.Tensor:
    """Embedding -> 1-layer LSTM (zero initial state, batch_first) -> final hidden state [B,H]
    (model.py:47-58).  Explicit recurrence, torch gate order i,f,g,o."""
    emb = p["text.wembedding.weight"][qst_idxs]            # [B,T,E]; row 0 is a learned row (no padding_idx)
    w_ih, w_hh = p["text.lstm.weight_ih_l0"], p["text.lstm.weight_hh_l0"]
    bias = p["text.lstm.bias_ih_l0"] + p["text.lstm.bias_hh_l0"]
    B, T, _ = emb.shape
    H = w_hh.shape[1]
    h = emb.new_zeros(B, H)
    c = emb.new_zeros(B, H)
    for t in range(T):
        gates = emb[:, t] @ w_ih.t() + h @ w_hh.t() + bias
        i, f, g, o = gates.split(H, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
    return h


# --------------------------------------------------------------------------------------
# relation layer -- literal formulation (reference model.py:104-162)
# --------------------------------------------------------------------------------------
def relation_pairs(x: torch.Tensor) -> torch.Tensor:
    """All n*n ordered pairs.  Row p = a*n + c of sample b is [x[b,c] | x[b,a]]
    (model.py:117-127: x_i repeats along dim 1, x_j along dim 2, cat on features)."""
    b, n, k = x.shape
    x_c = x.unsqueeze(1).expand(b, n, n, k)     # [b, a, c] -> x[b, c]
    x_a = x.unsqueeze(2).expand(b, n, n, k)     # [b, a, c] -> x[b, a]
    return torch.cat([x_c, x_a], dim=3).reshape(b * n * n, 2 * k)


def g_layer_params(p: Params, n_layers: int) -> List[Tuple[torch.Tensor, torch.Tensor]]:
    return [(p[f"rl.g_layers.{l}.weight"], p[f"rl.g_layers.{l}.bias"]) for l in range(n_layers)]


def g_mlp_dense(x: torch.Tensor, q: torch.Tensor, g_params, qinj: int,
                return_hidden: bool = False):
    """Literal g: pairs -> shared MLP with the question concatenated at layer ``qinj``
    (model.py:130-145) -> sum over all n*n pairs (model.py:151-152).  Returns x_g [B,G]."""
    b, n, _ = x.shape
    h = relation_pairs(x)
    hidden = []
    for l, (w, bias) in enumerate(g_params):
        if l == qinj:
            qrep = q.unsqueeze(1).expand(b, n * n, q.shape[1]).reshape(b * n * n, q.shape[1])
            h = torch.cat([h, qrep], dim=1)
        h = torch.relu(F.linear(h, w, bias))
        if return_hidden:
            hidden.append(h)
    x_g = h.reshape(b, n * n, h.shape[1]).sum(1)
    return (x_g, hidden) if return_hidden else x_g


def f_mlp(p: Params, x_g: torch.Tensor, dropout_p: float, training: bool,
          dropout_mask: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fc1 -> ReLU -> fc2 -> Dropout -> ReLU -> fc3 -> log_softmax (model.py:155-162; note the
    dropout sits BEFORE the ReLU).  ``dropout_mask`` (0/1, same shape as the fc2 output) makes
    train-mode results reproducible without sharing RNG streams: y = x * mask / (1-p)."""
    h = torch.relu(F.linear(x_g, p["rl.f_fc1.weight"], p["rl.f_fc1.bias"]))
    h = F.linear(h, p["rl.f_fc2.weight"], p["rl.f_fc2.bias"])
    if training and dropout_p > 0:
        if dropout_mask is None:
            h = F.dropout(h, dropout_p, True)
        else:
            h = h * dropout_mask.to(h.dtype) / (1.0 - dropout_p)
    h = torch.relu(h)
    h = F.linear(h, p["rl.f_fc3.weight"], p["rl.f_fc3.bias"])
    return F.log_softmax(h, dim=1)


def rn_forward(p: Params, hyp: dict, img: torch.Tensor, qst_idxs: torch.Tensor, training: bool = False,
               dropout_mask: Optional[torch.Tensor] = None, running_out: Optional[Params] = None,
               return_parts: bool = False):
    """Whole model (model.py:187-205).  img is [B,3,S,S] pixels, or [B,n,k] objects when
    hyp['state_description'] (model.py:188-189: no conv, no coords)."""
    if hyp["state_description"]:
        x = img
        feat = None
    else:
        feat = conv_features(p, img, training, running_out)
        x = objects_from_features(feat)
    q = question_embed(p, qst_idxs)
    g_params = g_layer_params(p, len(hyp["g_layers"]))
    x_g = g_mlp_dense(x, q, g_params, hyp["question_injection_position"])
    logp = f_mlp(p, x_g, hyp["dropout"], training, dropout_mask)
    if return_parts:
        return logp, dict(feat=feat, x=x, q=q, x_g=x_g)
    return logp


def training_step_loss(p: Params, hyp: dict, img, qst_idxs, label, dropout_mask=None) -> torch.Tensor:
    """Loss of one training step (train.py:40-41): mean NLL of the train-mode forward."""
    return F.nll_loss(rn_forward(p, hyp, img, qst_idxs, True, dropout_mask), label)


# --------------------------------------------------------------------------------------
# relation layer -- factorised formulation used by the CUDA kernels
# --------------------------------------------------------------------------------------
def split_g_weights(g_params, k: int, Q: int, qinj: int):
    """Split layer weights into the pieces the factorised form uses.
    Returns (W0c [G,k], W0a [G,k], Wh list for l>=1 [G,G], Wq [G,Q], biases list)."""
    w0 = g_params[0][0]
    W0c, W0a = w0[:, :k], w0[:, k:2 * k]
    Wh, biases = [], [g_params[0][1]]
    Wq = None
    if qinj == 0:
        Wq = w0[:, 2 * k:2 * k + Q]
    for l in range(1, len(g_params)):
        w, bias = g_params[l]
        gin = g_params[l - 1][0].shape[0]
        Wh.append(w[:, :gin])
        if l == qinj:
            Wq = w[:, gin:gin + Q]
        biases.append(bias)
    return W0c, W0a, Wh, Wq, biases


def g_mlp_factorised(x: torch.Tensor, q: torch.Tensor, g_params, qinj: int):
    """Exact algebraic restatement of g:
         U = X W0c^T, V = X W0a^T, beta_l = b_l (+ q Wq^T at l == qinj)
         Z1[a,c] = U[c] + V[a] + beta_0 ; H_l = relu(Z_l) ; Z_{l+1} = H_l Wh_l^T + beta_l
       Returns x_g and the saved intermediates (U, V, betas, H list, each H_l [B, n*n, G])."""
    b, n, k = x.shape
    Q = q.shape[1]
    W0c, W0a, Wh, Wq, biases = split_g_weights(g_params, k, Q, qinj)
    U = x @ W0c.t()                                  # [b,n,G]
    V = x @ W0a.t()
    betas = [bias.unsqueeze(0).expand(b, -1) for bias in biases]
    betas[qinj] = betas[qinj] + q @ Wq.t()           # [b,G] per-sample bias
    Z = U.unsqueeze(1) + V.unsqueeze(2) + betas[0][:, None, None, :]     # [b, a, c, G]
    H = [torch.relu(Z).reshape(b, n * n, -1)]
    for l, w in enumerate(Wh, start=1):
        Zl = H[-1] @ w.t() + betas[l][:, None, :]
        H.append(torch.relu(Zl))
    return H[-1].sum(1), dict(U=U, V=V, betas=betas, H=H)


def g_backward_factorised(x, q, g_params, qinj, saved, dxg):
    """Hand-written backward of g_mlp_factorised, given dxg [B,G].
    Returns dict(dx [B,n,k], dq [B,Q], dW list, db list) matching autograd of g_mlp_dense."""
    b, n, k = x.shape
    Q = q.shape[1]
    W0c, W0a, Wh, Wq, biases = split_g_weights(g_params, k, Q, qinj)
    H = saved["H"]
    L = len(g_params)
    dW = [None] * L
    db = [None] * L
    dq = torch.zeros_like(q)
    dZ = dxg[:, None, :] * (H[L - 1] > 0).to(dxg.dtype)            # [b, n*n, G]
    for l in range(L - 1, 0, -1):
        delta = dZ.sum(1)                                            # [b,G]
        dWh = torch.einsum("bpo,bpi->oi", dZ, H[l - 1])
        db[l] = delta.sum(0)
        if l == qinj:
            dW[l] = torch.cat([dWh, delta.t() @ q], dim=1)
            dq = dq + delta @ Wq
        else:
            dW[l] = dWh
        dZ = (dZ @ Wh[l - 1]) * (H[l - 1] > 0).to(dxg.dtype)
    dZ1 = dZ.reshape(b, n, n, -1)                                    # [b, a, c, G]
    dU = dZ1.sum(1)                                                  # over a -> [b, c, G]
    dV = dZ1.sum(2)                                                  # over c -> [b, a, G]
    delta0 = dU.sum(1)
    parts = [torch.einsum("bng,bnk->gk", dU, x), torch.einsum("bng,bnk->gk", dV, x)]
    if qinj == 0:
        parts.append(delta0.t() @ q)
        dq = dq + delta0 @ Wq
    dW[0] = torch.cat(parts, dim=1)
    db[0] = delta0.sum(0)
    dx = dU @ W0c + dV @ W0a
    return dict(dx=dx, dq=dq, dW=dW, db=db, dU=dU, dV=dV)


# --------------------------------------------------------------------------------------
# optimiser tail (reference train.py:45-48,330): clip_grad_norm(50) + Adam(weight_decay=1e-4)
# --------------------------------------------------------------------------------------
def clip_and_adam(params: List[torch.Tensor], grads: List[torch.Tensor], exp_avg, exp_avg_sq, step: int,
                  lr: float, clip_norm: float = 50.0, weight_decay: float = 1e-4,
                  betas=(0.9, 0.999), eps: float = 1e-8):
    """In-place restatement of torch.nn.utils.clip_grad_norm_ followed by torch.optim.Adam with
    L2 weight decay folded into the gradient (train.py:45-48, 330).  Returns the total norm."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).to(grads[0].dtype)
    coef = torch.clamp(clip_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    for w, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = g * coef + weight_decay * w
        m.mul_(b1).add_(g, alpha=1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / math.sqrt(1 - b2 ** step)) + eps
        w.addcdiv_(m, denom, value=-lr / (1 - b1 ** step))
    return total


# --------------------------------------------------------------------------------------
# synthetic inputs shaped like the reference's data contract
# --------------------------------------------------------------------------------------
def structured_images(B: int, side: int, seed: int) -> torch.Tensor:
    """Flat background + a few coloured rectangles, in [0,1) like ToTensor (train.py:186).
    Trained checkpoints are extremely ReLU-sparse on uniform noise, so parity on trained
    weights uses these."""
    import numpy as np

    rng = np.random.default_rng(seed)
    img = np.empty((B, 3, side, side), dtype=np.float32)
    for b in range(B):
        img[b] = rng.uniform(0.3, 0.6, size=(3, 1, 1))
        for _ in range(int(rng.integers(3, 8))):
            w, h = rng.integers(side // 10, side // 4, size=2)
            x0, y0 = rng.integers(0, side - w), rng.integers(0, side - h)
            img[b, :, y0:y0 + h, x0:x0 + w] = rng.uniform(0.0, 1.0, size=(3, 1, 1))
        img[b] += rng.normal(0, 0.01, size=(3, side, side)).astype(np.float32)
    return torch.from_numpy(np.clip(img, 0.0, 0.999))


def uniform_images(B: int, side: int, seed: int) -> torch.Tensor:
    import numpy as np

    return torch.from_numpy(np.random.default_rng(seed).random((B, 3, side, side), dtype=np.float32))


def state_descriptions(B: int, seed: int, n_real: int = 10, n_pad: int = 12) -> torch.Tensor:
    """[B,12,7] object rows [x,y,z,color,material,shape,size], zero rows as padding
    (utils.py:101-107; clevr_dataset_connector.py:109-117)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    o = np.zeros((B, n_pad, 7), dtype=np.float32)
    o[:, :n_real, 0:2] = rng.uniform(-3, 3, size=(B, n_real, 2))
    o[:, :n_real, 2] = rng.choice([0.35, 0.7], size=(B, n_real))
    o[:, :n_real, 3] = rng.integers(1, 9, size=(B, n_real))
    o[:, :n_real, 4] = rng.integers(1, 3, size=(B, n_real))
    o[:, :n_real, 5] = rng.integers(1, 4, size=(B, n_real))
    o[:, :n_real, 6] = rng.integers(1, 3, size=(B, n_real))
    return torch.from_numpy(o)


def questions(B: int, T: int, qdict_size: int, seed: int, left_pad: int = 0) -> torch.Tensor:
    """[B,T] int64 token ids in 1..qdict_size, optionally with ``left_pad`` zeros on the left
    (reversed questions are left-padded with 0: utils.py:138-141)."""
    import numpy as np

    rng = np.random.default_rng(seed)
    q = rng.integers(1, qdict_size + 1, size=(B, T)).astype(np.int64)
    if left_pad:
        q[:, :left_pad] = 0
    return torch.from_numpy(q)


def labels(B: int, adict_size: int, seed: int) -> torch.Tensor:
    import numpy as np

    return torch.from_numpy(np.random.default_rng(seed).integers(0, adict_size, size=(B,)).astype(np.int64))


def rel_err(a: torch.Tensor, ref: torch.Tensor) -> float:
    """The tolerance metric used throughout: max|a-ref| / max|ref|."""
    denom = float(ref.abs().max())
    return float((a.double() - ref.double()).abs().max()) / (denom if denom > 0 else 1.0)
