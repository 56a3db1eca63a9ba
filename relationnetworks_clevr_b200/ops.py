"""torch.autograd bindings of the C-ABI entry points (one Function per reference span).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; all arithmetic of
these ops happens in librn_b200.so.  Every op requires CUDA tensors and raises otherwise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, Tuple

import torch

from . import _lib
from ._lib import PRECISION, AdamCfg, ConvCfg, ConvGrads, ConvLayer, FCfg, LstmCfg, RelationCfg, check, lib, ptr_array

_scratch: Dict[Tuple[int, str], torch.Tensor] = {}


def _scratch_bytes(device: torch.device, tag: str, nbytes: int) -> torch.Tensor:
    """Grow-only per-device scratch (dead between calls; all use is ordered on the current stream)."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), tag)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)
        _scratch[key] = buf
    return buf


# ---- side stream for work that overlaps the kernels of this library (model.RN runs the question encoder on it) ----
_side_streams: Dict[int, torch.cuda.Stream] = {}


def side_stream(device: torch.device) -> torch.cuda.Stream:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        st = torch.cuda.Stream(device=idx)
        _side_streams[idx] = st
    return st


# ---- auxiliary stream for the question encoder -----------------------------------------------------------------
# The LSTM kernels are short, latency-bound launches that depend only on the tokens and their own parameters (forward)
# or on dq (backward); the conv stack next to them is a chain of kernels that never fills the machine at small batches.
# They run concurrently: the encoder's C-ABI calls are issued on an auxiliary stream that waits on a FORK POINT (an event
# recorded on the main stream at the start of RN.forward / at the end of the relation backward, i.e. independent of the
# order in which autograd happens to call the two backward nodes) and the main stream JOINS before the first consumer
# (the relation forward; the optimiser).  No torch stream context is switched, so every tensor is allocated from -- and
# returns to -- the main stream's pool, and under CUDA-graph capture the fork / join pair is captured as a parallel branch.
_aux: Dict[int, dict] = {}


def _aux_state(device: torch.device) -> dict:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _aux.get(idx)
    if st is None:
        st = {"stream": torch.cuda.Stream(device=idx), "fork": None, "join": None, "pending": False, "keep": []}
        _aux[idx] = st
    return st


def fork_point(device: torch.device) -> None:
    """Work issued on the auxiliary stream from now on may start once everything enqueued on the main stream so far is done."""
    st = _aux_state(device)
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(device))
    st["fork"] = ev


def _aux_begin(device: torch.device) -> int:
    st = _aux_state(device)
    if st["fork"] is None:
        fork_point(device)
    st["stream"].wait_event(st["fork"])
    return st["stream"].cuda_stream


def _aux_end(device: torch.device, defer_join: bool, keep=()) -> None:
    """`keep`: tensors the auxiliary-stream kernels touch.  They are held until the join, so the caching allocator (which
    only knows the main stream) cannot hand their memory to main-stream work that would run concurrently."""
    st = _aux_state(device)
    ev = torch.cuda.Event()
    ev.record(st["stream"])
    st["join"], st["pending"], st["fork"] = ev, True, None
    st["keep"].extend(keep)
    if not defer_join:
        join_aux(device)


def join_aux(device: torch.device) -> None:
    """The main stream waits for everything issued on the auxiliary stream (no-op when nothing is pending)."""
    st = _aux.get(device.index if device.index is not None else torch.cuda.current_device())
    if st is not None and st["pending"]:
        torch.cuda.current_stream(device).wait_event(st["join"])
        st["pending"] = False
        st["keep"].clear()


def _mark_for_side_consumers(t: torch.Tensor) -> None:
    """A gradient produced here may be consumed by a backward node running on the side stream: keep the caching
    allocator from recycling it before that stream is done with it."""
    st = _side_streams.get(t.device.index if t.device.index is not None else torch.cuda.current_device())
    if st is not None:
        t.record_stream(st)


# ---- optional CUDA-event timers around the relation launches (bench.py's roofline leg) ----------
_timers_on = False
_timer_events: Dict[str, list] = {}


def timers_enable(flag: bool) -> None:
    global _timers_on
    _timers_on = bool(flag)
    if flag:
        _timer_events.clear()


class _Timed:
    """Records a CUDA event pair on the current stream around a C-ABI call when timers are enabled."""

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if _timers_on:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e1 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if _timers_on:
            self.e1.record()
            _timer_events.setdefault(self.name, []).append((self.e0, self.e1))
        return False


def timers_collect() -> Dict[str, list]:
    """Milliseconds per recorded call (synchronises)."""
    torch.cuda.synchronize()
    return {k: [a.elapsed_time(b) for a, b in v] for k, v in _timer_events.items()}


# ---- gradient sink: parameter gradients written straight into the optimiser's flat gradient buffer ----------------
# trainer.FlatClipAdam keeps every parameter as a view of ONE flat fp32 buffer.  While a sink is registered, the backward
# of each op writes the gradient of such a parameter directly into the matching slice of the flat GRADIENT buffer and
# returns None for it to autograd (param.grad stays None): no per-parameter gradient tensors, no torch.cat, no zero
# fills -- every slice is overwritten by exactly one kernel per step.  Without a sink the ops return ordinary tensors.
_sink = None      # (flat_params, flat_grads)


def set_grad_sink(flat_params: torch.Tensor, flat_grads: torch.Tensor) -> None:
    global _sink
    if flat_params.numel() != flat_grads.numel() or flat_params.dtype != torch.float32 or flat_grads.dtype != torch.float32:
        raise RuntimeError("gradient sink: flat parameter and gradient buffers must be fp32 and equally long")
    _sink = (flat_params, flat_grads)


def clear_grad_sink() -> None:
    global _sink
    _sink = None


def _sink_view(p: torch.Tensor):
    if _sink is None or not p.is_contiguous() or p.dtype != torch.float32:
        return None
    flat, grads = _sink
    if p.device != flat.device:
        return None
    off = p.data_ptr() - flat.data_ptr()
    if off < 0 or off % 4 or off // 4 + p.numel() > flat.numel():
        return None
    return grads[off // 4: off // 4 + p.numel()].view(p.shape)


def _grad_outputs(params, like):
    """One output tensor per parameter: its slice of the flat gradient buffer when a sink covers it, else a new tensor."""
    out = []
    for p, l in zip(params, like):
        v = _sink_view(p)
        out.append(v if v is not None else torch.empty_like(l))
    return out


def _grad_returns(params, grads):
    return [None if _sink_view(p) is not None else g for p, g in zip(params, grads)]


def _require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("relationnetworks_clevr_b200 ops run on CUDA (sm_100a) tensors only; "
                               "there is no CPU path (use oracle/ for CPU checks in tests)")


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _on_tensor_device(fn):
    """Run an autograd.Function forward / backward with the CUDA device of its first tensor argument current, so the
    stream, the scratch buffers and the SM count all belong to the device the data lives on (a model moved with
    ``.cuda(1)`` without ``torch.cuda.set_device(1)``, autograd worker threads)."""
    import functools

    @functools.wraps(fn)
    def wrapper(ctx, *args):
        dev = next((a.device for a in args if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            saved = getattr(ctx, "saved_tensors", ())
            dev = next((a.device for a in saved if isinstance(a, torch.Tensor) and a.is_cuda), None)
        if dev is None:
            return fn(ctx, *args)
        with torch.cuda.device(dev):
            return fn(ctx, *args)

    return wrapper


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


# rn_relation_cfg.flags (include/rn_b200.h RN_REL_FLAG_*); diagnostics may set RN_B200_REL_FLAGS
REL_FLAG_FWD_2PASS = 1
REL_FLAG_DGRAD_2PASS = 2
relation_flags = int(os.environ.get("RN_B200_REL_FLAGS", "0"))
# rn_conv_cfg.flags (RN_CONV_FLAG_*): 1 = fp32 SIMT convolutions everywhere, 2 = tensor-core forward as well (tests, A/B timing)
conv_flags = int(os.environ.get("RN_B200_CONV_FLAGS", "0"))
# The tensor-core g-MLP keeps activations as fp16 (hi + lo): a hidden activation above 65504 overflows to inf, which the
# fp32 reference would survive.  Trained and seeded CLEVR models sit 3 orders of magnitude below that, so the check is
# opt-in (it costs a device synchronisation per forward): RN_B200_CHECK_FINITE=1 raises instead of returning inf / NaN.
check_finite = os.environ.get("RN_B200_CHECK_FINITE", "0") == "1"


def relation_cfg(B, n, k, Q, G, L, qinj, precision: str, training: bool, flags=None) -> RelationCfg:
    return RelationCfg(B, n, k, Q, G, L, qinj, PRECISION[precision], int(training),
                       relation_flags if flags is None else flags)


def tc_supported(n: int, G: int, L: int, k: int = 26, Q: int = 128, qinj: int = 0) -> bool:
    cfg = relation_cfg(1, n, k, Q, G, L, qinj, "parity", False)
    return bool(lib().rn_relation_tc_supported(C.byref(cfg)))


class RelationFunction(torch.autograd.Function):
    """x_g = sum over all n*n ordered pairs of g([x_c | x_a | q])  (reference model.py:104-152).

    forward(x [B,n,k], q [B,Q], qinj, precision, w0, b0, ..., w_{L-1}, b_{L-1}) -> x_g [B,G]
    """

    @staticmethod
    @_on_tensor_device
    def forward(ctx, x, q, qinj, precision, *wb):
        _require_cuda(x, q, *wb)
        L = len(wb) // 2
        ws = [_f32c(t) for t in wb[0::2]]
        bs = [_f32c(t) for t in wb[1::2]]
        x_, q_ = _f32c(x), _f32c(q)
        B, n, k = x_.shape
        Q, G = q_.shape[1], ws[0].shape[0]
        training = any(ctx.needs_input_grad)      # forward runs with grad mode off; this is the autograd signal
        cfg = relation_cfg(B, n, k, Q, G, L, qinj, precision, training)
        for l, w in enumerate(ws):
            fan = (2 * k if l == 0 else G) + (Q if l == qinj else 0)
            if tuple(w.shape) != (G, fan):
                raise RuntimeError(f"g layer {l} weight has shape {tuple(w.shape)}, expected {(G, fan)}")
        sb, cb = C.c_size_t(), C.c_size_t()
        check(lib().rn_relation_workspace(C.byref(cfg), C.byref(sb), C.byref(cb)), "rn_relation_workspace")
        saved = torch.empty(max(sb.value, 256), dtype=torch.uint8, device=x.device)
        scratch = _scratch_bytes(x.device, "relation", cb.value)
        xg = torch.empty(B, G, dtype=torch.float32, device=x.device)
        with _Timed("relation_fwd"):
            check(lib().rn_relation_fwd(C.byref(cfg), x_.data_ptr(), q_.data_ptr(), ptr_array(ws), ptr_array(bs),
                                        xg.data_ptr(), saved.data_ptr(), scratch.data_ptr(), _stream()),
                  "rn_relation_fwd")
        if check_finite and precision != "fp32" and not torch.cuda.is_current_stream_capturing() \
                and not bool(torch.isfinite(xg).all()):
            raise RuntimeError("rn_relation_fwd: non-finite x_g -- a g-layer activation left the fp16 range of the "
                               "tensor-core path (|h| > 65504) or the inputs are non-finite; use precision='fp32'")
        if training:
            ctx.cfg = cfg
            ctx.saved_buf = saved
            ctx.L = L
            ctx.params = (wb[0::2], wb[1::2])
            ctx.save_for_backward(x_, q_, *ws)
        return xg

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dxg):
        x_, q_, *ws = ctx.saved_tensors
        cfg, L = ctx.cfg, ctx.L
        dxg_ = _f32c(dxg)
        dx = torch.empty_like(x_)
        dq = torch.empty_like(q_)
        pw, pb = ctx.params
        dws = _grad_outputs(pw, ws)
        dbs = _grad_outputs(pb, [torch.empty(w.shape[0], dtype=torch.float32, device=x_.device) for w in ws])
        sb, cb = C.c_size_t(), C.c_size_t()
        check(lib().rn_relation_workspace(C.byref(cfg), C.byref(sb), C.byref(cb)), "rn_relation_workspace")
        scratch = _scratch_bytes(x_.device, "relation", cb.value)
        with _Timed("relation_bwd"):
            check(lib().rn_relation_bwd(C.byref(cfg), dxg_.data_ptr(), x_.data_ptr(), q_.data_ptr(), ptr_array(ws),
                                        ctx.saved_buf.data_ptr(), dx.data_ptr(), dq.data_ptr(), ptr_array(dws),
                                        ptr_array(dbs), scratch.data_ptr(), _stream()), "rn_relation_bwd")
        _mark_for_side_consumers(dq)
        fork_point(x_.device)          # dq and dx exist: the question-encoder backward may run next to the conv backward
        grads = []
        for dw, db in zip(_grad_returns(pw, dws), _grad_returns(pb, dbs)):
            grads += [dw, db]
        return (dx, dq, None, None, *grads)


def relation_forward_materialised(x, q, qinj, *wb):
    """Eval-time forward of the relation op on the fp32 kernels that KEEPS every g-layer activation in device memory:
    returns (x_g [B,G], [H_1 .. H_L]) with H_{l+1} = relu(g layer l) as [B*n*n, G] views of one buffer.  This is what
    forward hooks on ``rl.g_layers[i]`` (reference extract.py:43-47) need to see; the training / inference paths never
    materialise these tensors (2.7 GB each at B = 640)."""
    _require_cuda(x, q, *wb)
    with torch.no_grad(), torch.cuda.device(x.device):
        ws = [_f32c(t) for t in wb[0::2]]
        bs = [_f32c(t) for t in wb[1::2]]
        x_, q_ = _f32c(x), _f32c(q)
        B, n, k = x_.shape
        Q, G, L = q_.shape[1], ws[0].shape[0], len(ws)
        cfg = relation_cfg(B, n, k, Q, G, L, qinj, "fp32", True)
        sb, cb = C.c_size_t(), C.c_size_t()
        check(lib().rn_relation_workspace(C.byref(cfg), C.byref(sb), C.byref(cb)), "rn_relation_workspace")
        saved = torch.empty(max(sb.value, 256), dtype=torch.uint8, device=x.device)
        scratch = _scratch_bytes(x.device, "relation", cb.value)
        xg = torch.empty(B, G, dtype=torch.float32, device=x.device)
        check(lib().rn_relation_fwd(C.byref(cfg), x_.data_ptr(), q_.data_ptr(), ptr_array(ws), ptr_array(bs), xg.data_ptr(),
                                    saved.data_ptr(), scratch.data_ptr(), _stream()), "rn_relation_fwd")
        acts = []
        base = saved.data_ptr()
        for l in range(L):
            ptr = C.c_void_p()
            check(lib().rn_relation_activation(C.byref(cfg), base, l, C.byref(ptr)), "rn_relation_activation")
            off = ptr.value - base
            acts.append(saved[off: off + B * n * n * G * 4].view(torch.float32).view(B * n * n, G))
        return xg, acts


def extract_stats(z: torch.Tensor, B: int, width: int):
    """(max, mean) over the rows of each sample of the L2-normalised rows of z [B*P, ld], first `width` columns
    (the aggregation of reference extract.py:63-74), one pass over z."""
    _require_cuda(z)
    z = _f32c(z)
    P = z.shape[0] // B
    with torch.cuda.device(z.device):
        maxf = torch.empty(B, width, dtype=torch.float32, device=z.device)
        avgf = torch.empty(B, width, dtype=torch.float32, device=z.device)
        scratch = _scratch_bytes(z.device, "extract", 4 * B * 32 * 2 * width)
        check(lib().rn_extract_stats(z.data_ptr(), B, P, z.shape[1], width, maxf.data_ptr(), avgf.data_ptr(), scratch.data_ptr(),
                                     _stream()), "rn_extract_stats")
    return maxf, avgf


class FHeadFunction(torch.autograd.Function):
    """log_softmax(fc3(relu(dropout(fc2(relu(fc1(x_g)))))))  (reference model.py:155-162).

    ``drop_mask`` is a uint8 [B,F2] keep-mask drawn by the caller from torch's RNG (or None)."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, xg, w1, b1, w2, b2, w3, b3, drop_mask, keep_scale):
        _require_cuda(xg, w1, b1, w2, b2, w3, b3, drop_mask)
        t = [_f32c(v) for v in (xg, w1, b1, w2, b2, w3, b3)]
        B, G = t[0].shape
        F1, F2, A = t[1].shape[0], t[3].shape[0], t[5].shape[0]
        cfg = FCfg(B, G, F1, F2, A, float(keep_scale))
        logp = torch.empty(B, A, dtype=torch.float32, device=xg.device)
        saved = torch.empty(B * (F1 + F2), dtype=torch.float32, device=xg.device)
        mask_ptr = drop_mask.data_ptr() if drop_mask is not None else None
        if drop_mask is not None and (drop_mask.dtype != torch.uint8 or tuple(drop_mask.shape) != (B, F2)):
            raise RuntimeError("drop_mask must be uint8 [B, F2]")
        with _Timed("f_fwd"):
            check(lib().rn_f_fwd(C.byref(cfg), *[v.data_ptr() for v in t], mask_ptr, logp.data_ptr(), saved.data_ptr(),
                                 _stream()), "rn_f_fwd")
        ctx.cfg = cfg
        ctx.has_mask = drop_mask is not None
        ctx.params = (w1, b1, w2, b2, w3, b3)
        ctx.save_for_backward(logp, saved, t[0], t[1], t[3], t[5], drop_mask if drop_mask is not None else logp)
        return logp

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dlogp):
        logp, saved, xg, w1, w2, w3, mask = ctx.saved_tensors
        cfg = ctx.cfg
        dev = xg.device
        dl = _f32c(dlogp)
        dxg = torch.empty_like(xg)
        like = [w1, torch.empty(cfg.F1, dtype=torch.float32, device=dev), w2, torch.empty(cfg.F2, dtype=torch.float32, device=dev),
                w3, torch.empty(cfg.A, dtype=torch.float32, device=dev)]
        dw1, db1, dw2, db2, dw3, db3 = _grad_outputs(ctx.params, like)
        scratch = torch.empty(cfg.B * (cfg.A + cfg.F2 + cfg.F1), dtype=torch.float32, device=dev)
        check(lib().rn_f_bwd(C.byref(cfg), dl.data_ptr(), logp.data_ptr(), xg.data_ptr(), w1.data_ptr(), w2.data_ptr(),
                             w3.data_ptr(), mask.data_ptr() if ctx.has_mask else None, saved.data_ptr(),
                             dxg.data_ptr(), dw1.data_ptr(), db1.data_ptr(), dw2.data_ptr(), db2.data_ptr(),
                             dw3.data_ptr(), db3.data_ptr(), scratch.data_ptr(), _stream()), "rn_f_bwd")
        return (dxg, *_grad_returns(ctx.params, [dw1, db1, dw2, db2, dw3, db3]), None, None)


# The auxiliary stream pays when the batch is small (at 80 questions per GPU the conv kernels cover a fraction of the SMs:
# measured 1.53 vs X ms per step); at 640 per GPU the encoder's 144 register-heavy CTAs take SMs away from the conv stack
# and the step gets 2 % slower (8.11 vs 7.93 ms), so it is used below a batch threshold only.
use_aux_stream = os.environ.get("RN_B200_TEXT_STREAM", "1") != "0"
aux_stream_max_batch = int(os.environ.get("RN_B200_TEXT_STREAM_MAX_BATCH", "256"))
_defer_forward_join = False      # set by RN.forward around its text call: it joins itself, before the relation op


class defer_text_join:
    """Context manager used by RN.forward: inside it the question encoder's forward leaves its join to the caller."""

    def __enter__(self):
        global _defer_forward_join
        self.prev, _defer_forward_join = _defer_forward_join, True

    def __exit__(self, *exc):
        global _defer_forward_join
        _defer_forward_join = self.prev
        return False


def lstm_supported(B: int, T: int, V: int, E: int, H: int) -> bool:
    cfg = LstmCfg(B, T, V, E, H, 0)
    return bool(lib().rn_lstm_supported(C.byref(cfg)))


class QuestionEncoderFunction(torch.autograd.Function):
    """q = last hidden state of LSTM(Embedding(tokens)) with a zero initial state (reference model.py:47-58).

    forward(tokens [B,T] int64, emb [V,E], w_ih [4H,E], w_hh [4H,H], b_ih [4H], b_hh [4H]) -> q [B,H]"""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, tokens, emb, w_ih, w_hh, b_ih, b_hh):
        _require_cuda(tokens, emb, w_ih, w_hh, b_ih, b_hh)
        if tokens.dtype != torch.int64:
            raise RuntimeError("question tokens must be int64 (as for nn.Embedding)")
        tok = tokens.contiguous()
        ps = [_f32c(t) for t in (emb, w_ih, w_hh, b_ih, b_hh)]
        B, T = tok.shape
        V, E = ps[0].shape
        H = ps[2].shape[1]
        training = any(ctx.needs_input_grad)
        cfg = LstmCfg(B, T, V, E, H, int(training))
        sf, cf = C.c_size_t(), C.c_size_t()
        check(lib().rn_lstm_workspace(C.byref(cfg), C.byref(sf), C.byref(cf)), "rn_lstm_workspace")
        saved = torch.empty(sf.value, dtype=torch.float32, device=tok.device)
        q = torch.empty(B, H, dtype=torch.float32, device=tok.device)
        aux = use_aux_stream and B <= aux_stream_max_batch
        stream = _aux_begin(tok.device) if aux else _stream()
        with _Timed("lstm_fwd"):
            check(lib().rn_lstm_fwd(C.byref(cfg), tok.data_ptr(), *[p.data_ptr() for p in ps], q.data_ptr(), saved.data_ptr(),
                                    stream), "rn_lstm_fwd")
        if aux:      # inside RN.forward the join is RN's (before the relation op reads q); standalone callers join here
            _aux_end(tok.device, defer_join=_defer_forward_join, keep=(tok, saved, q, *ps))
        ctx.aux = aux
        if training:
            ctx.cfg = cfg
            ctx.scratch_floats = cf.value
            ctx.params = (emb, w_ih, w_hh, b_ih, b_hh)
            ctx.save_for_backward(tok, saved, *ps)
        return q

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dq):
        tok, saved, *ps = ctx.saved_tensors
        dq_ = _f32c(dq)
        grads = _grad_outputs(ctx.params, ps)
        scratch = _scratch_bytes(tok.device, "lstm", ctx.scratch_floats * 4)
        # The auxiliary stream is only safe when this call allocates NOTHING: a block the caching allocator hands out here
        # may have just been released by the conv backward, whose kernels are still running on the main stream (the fork
        # point precedes them) -- measured: a fresh gradient tensor landed on the saved input image and corrupted
        # dW(conv1).  With the optimiser's gradient sink every output is a slice of the flat gradient buffer.
        aux = ctx.aux and _sink is not None and all(_sink_view(p) is not None for p in ctx.params) and dq_.data_ptr() == dq.data_ptr()
        stream = _aux_begin(tok.device) if aux else _stream()
        with _Timed("lstm_bwd"):
            check(lib().rn_lstm_bwd(C.byref(ctx.cfg), tok.data_ptr(), ps[0].data_ptr(), ps[1].data_ptr(), ps[2].data_ptr(),
                                    dq_.data_ptr(), saved.data_ptr(), *[g.data_ptr() for g in grads], scratch.data_ptr(),
                                    stream), "rn_lstm_bwd")
        if aux:      # the join is FlatClipAdam.step()'s
            _aux_end(tok.device, defer_join=True, keep=(tok, saved, dq_, scratch, *ps, *grads))
        return (None, *_grad_returns(ctx.params, grads))


def _conv_layers(params, running) -> C.Array:
    arr = (ConvLayer * 4)()
    for l in range(4):
        w, b, g, beta = params[4 * l: 4 * l + 4]
        rm, rv = running[2 * l: 2 * l + 2]
        arr[l] = ConvLayer(w.data_ptr(), b.data_ptr(), g.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr())
    return arr


class ConvObjectsFunction(torch.autograd.Function):
    """objects [B, d*d, 26] = coords-augmented output of 4 x [conv3x3 s2 p1 -> BatchNorm -> ReLU]
    (reference model.py:22-36 and 192-201).

    forward(img, training, eps, momentum, running(8 tensors, updated in place), 16 params)
    params per layer: conv weight, conv bias, bn weight, bn bias."""

    @staticmethod
    @_on_tensor_device
    def forward(ctx, img, training, eps, momentum, running, *params):
        _require_cuda(img, *params, *running)
        img_u8 = img.dtype == torch.uint8        # raw pixels: converted (u / 255) inside the first conv layer's staging
        img_ = img.detach().contiguous() if img_u8 else _f32c(img)
        ps = [_f32c(p) for p in params]
        for r in running:
            if r.dtype != torch.float32 or not r.is_contiguous():
                raise RuntimeError("BatchNorm running statistics must be contiguous fp32")
        B, cin, side, side2 = img_.shape
        if cin != 3 or side != side2:
            raise RuntimeError(f"expected [B,3,S,S] images, got {tuple(img_.shape)}")
        cfg = ConvCfg(B, side, int(training), float(eps), float(momentum), int(img_u8), conv_flags)
        sf, cf = C.c_size_t(), C.c_size_t()
        check(lib().rn_conv_workspace(C.byref(cfg), C.byref(sf), C.byref(cf)), "rn_conv_workspace")
        saved = torch.empty(sf.value, dtype=torch.float32, device=img.device)
        scratch = _scratch_bytes(img.device, "conv", cf.value * 4)
        d = side // 16
        objects = torch.empty(B, d * d, 26, dtype=torch.float32, device=img.device)
        layers = _conv_layers(ps, running)
        with _Timed("conv_fwd"):
            check(lib().rn_conv_fwd(C.byref(cfg), img_.data_ptr(), layers, objects.data_ptr(), saved.data_ptr(),
                                    scratch.data_ptr(), _stream()), "rn_conv_fwd")
        if any(ctx.needs_input_grad):
            ctx.cfg = cfg
            ctx.saved_buf = saved
            ctx.running = running
            ctx.scratch_floats = cf.value
            ctx.params = params
            ctx.save_for_backward(img_, *ps)
        return objects

    @staticmethod
    @_on_tensor_device
    def backward(ctx, dobjects):
        img_, *ps = ctx.saved_tensors
        cfg = ctx.cfg
        dobj = _f32c(dobjects)
        grads = _grad_outputs(ctx.params, ps)
        garr = (ConvGrads * 4)()
        for l in range(4):
            garr[l] = ConvGrads(*[g.data_ptr() for g in grads[4 * l: 4 * l + 4]])
        layers = _conv_layers(ps, ctx.running)
        scratch = _scratch_bytes(img_.device, "conv", ctx.scratch_floats * 4)
        with _Timed("conv_bwd"):
            check(lib().rn_conv_bwd(C.byref(cfg), img_.data_ptr(), dobj.data_ptr(), layers, ctx.saved_buf.data_ptr(),
                                    garr, scratch.data_ptr(), _stream()), "rn_conv_bwd")
        return (None, None, None, None, None, *_grad_returns(ctx.params, grads))


def clip_adam_(params: torch.Tensor, grads: torch.Tensor, exp_avg: torch.Tensor, exp_avg_sq: torch.Tensor, step: int,
               lr: float, clip_norm: float = 50.0, weight_decay: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
               grad_scale: float = 1.0, step_dev: torch.Tensor = None, lr_dev: torch.Tensor = None,
               total_out: torch.Tensor = None) -> torch.Tensor:
    """Fused clip_grad_norm + Adam(weight_decay) on flat fp32 buffers (reference train.py:45-48, 330).
    Returns the (pre-clip) total gradient norm as a 1-element device tensor.  `step_dev` (int32[1]) / `lr_dev` (fp32[1]):
    device-resident step counter (incremented by the call) and learning rate, for CUDA-graph replays."""
    _require_cuda(params, grads, exp_avg, exp_avg_sq, step_dev, lr_dev)
    n = params.numel()
    cfg = AdamCfg(n, lr, betas[0], betas[1], eps, weight_decay, clip_norm if clip_norm else 0.0, grad_scale, step)
    norm_scratch = _scratch_bytes(params.device, "adam", 4 * 1032)
    total = total_out if total_out is not None else torch.empty(1, dtype=torch.float32, device=params.device)
    check(lib().rn_clip_adam(C.byref(cfg), params.data_ptr(), grads.data_ptr(), exp_avg.data_ptr(),
                             exp_avg_sq.data_ptr(), norm_scratch.data_ptr(), total.data_ptr(),
                             step_dev.data_ptr() if step_dev is not None else None,
                             lr_dev.data_ptr() if lr_dev is not None else None, _stream()), "rn_clip_adam")
    return total
