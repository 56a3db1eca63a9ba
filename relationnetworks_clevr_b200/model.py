"""Drop-in mirror of the reference's ``model.py`` class surface on top of the sm_100a kernels.

Same class names, constructor arguments, attribute names and ``state_dict`` keys as the reference
(``conv.conv{1-4}``, ``conv.batchNorm{1-4}``, ``text.wembedding``, ``text.lstm``, ``rl.g_layers.{i}``,
``rl.f_fc{1,2,3}``), so both shipped checkpoints load with ``strict=True`` (minus ``module.``) and
``train.py --resume`` keeps working.  The nn.Modules only *hold parameters*; the arithmetic of conv,
g, the pair-sum and f runs in librn_b200.so through ``ops``.  The LSTM question encoder stays in
PyTorch (BASELINE.json north_star).

Reference: model.py:9-36 (ConvInputModel), :39-58 (QuestionEmbedModel), :60-78 (RelationalLayerBase),
:81-162 (RelationalLayer), :164-223 (RN).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


class ConvInputModel(nn.Module):
    """4 x [Conv2d(3x3, stride 2, pad 1) -> BatchNorm2d(24) -> ReLU] (reference model.py:9-36).

    ``forward(img)`` returns the reference's [B,24,d,d] feature map; ``objects(img)`` returns the
    coords-augmented [B, d*d, 26] object tensor RN consumes (model.py:192-201), straight from the kernel."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 24, 3, stride=2, padding=1)
        self.batchNorm1 = nn.BatchNorm2d(24)
        self.conv2 = nn.Conv2d(24, 24, 3, stride=2, padding=1)
        self.batchNorm2 = nn.BatchNorm2d(24)
        self.conv3 = nn.Conv2d(24, 24, 3, stride=2, padding=1)
        self.batchNorm3 = nn.BatchNorm2d(24)
        self.conv4 = nn.Conv2d(24, 24, 3, stride=2, padding=1)
        self.batchNorm4 = nn.BatchNorm2d(24)

    def _layers(self):
        return ((self.conv1, self.batchNorm1), (self.conv2, self.batchNorm2),
                (self.conv3, self.batchNorm3), (self.conv4, self.batchNorm4))

    def objects(self, img: torch.Tensor) -> torch.Tensor:
        params, running = [], []
        eps, momentum = self.batchNorm1.eps, self.batchNorm1.momentum
        for conv, bn in self._layers():
            params += [conv.weight, conv.bias, bn.weight, bn.bias]
            running += [bn.running_mean, bn.running_var]
        if self.training:       # nn.BatchNorm2d bookkeeping (unused by the arithmetic: momentum is a constant); one launch
            torch._foreach_add_([bn.num_batches_tracked for _, bn in self._layers() if bn.num_batches_tracked is not None], 1)
        return ops.ConvObjectsFunction.apply(img, self.training, eps, momentum, running, *params)

    def forward(self, img: torch.Tensor) -> torch.Tensor:
        obj = self.objects(img)                       # [B, d*d, 26]
        self.last_objects = obj                       # RN.forward reuses it when it goes through this module (hooks on `conv`)
        b, n, _ = obj.shape
        d = int(round(n ** 0.5))
        return obj[:, :, :24].permute(0, 2, 1).reshape(b, 24, d, d)


class QuestionEmbedModel(nn.Module):
    """Embedding -> LSTM -> last hidden state (reference model.py:39-58).  Stays in PyTorch."""

    def __init__(self, in_size, embed=32, hidden=128):
        super().__init__()
        self.wembedding = nn.Embedding(in_size + 1, embed)
        self.lstm = nn.LSTM(embed, hidden, batch_first=True)
        self.hidden = hidden
        self.use_kernel = os.environ.get("RN_B200_LSTM_KERNEL", "1") != "0"

    def forward(self, question: torch.Tensor) -> torch.Tensor:
        if self.use_kernel and question.is_cuda and question.dim() == 2 and ops.lstm_supported(
                question.shape[0], question.shape[1], self.wembedding.num_embeddings, self.wembedding.embedding_dim, self.hidden):
            # one persistent launch for all T steps (csrc/lstm.cu); the nn.LSTM module only holds the parameters
            return ops.QuestionEncoderFunction.apply(question, self.wembedding.weight, self.lstm.weight_ih_l0,
                                                     self.lstm.weight_hh_l0, self.lstm.bias_ih_l0, self.lstm.bias_hh_l0)
        # other shapes (hidden size 256 of the state-description configs, T > 64): PyTorch / cuDNN, fp32 (TF32 off, see __init__)
        wembed = self.wembedding(question)
        self.lstm.flatten_parameters()
        _, hidden = self.lstm(wembed)
        return hidden[0][0]


class RelationalLayerBase(nn.Module):
    """Owns the f-MLP parameters (reference model.py:60-78)."""

    def __init__(self, in_size, out_size, qst_size, hyp):
        super().__init__()
        self.f_fc1 = nn.Linear(hyp["g_layers"][-1], hyp["f_fc1"])
        self.f_fc2 = nn.Linear(hyp["f_fc1"], hyp["f_fc2"])
        self.f_fc3 = nn.Linear(hyp["f_fc2"], out_size)
        self.dropout = nn.Dropout(p=hyp["dropout"])
        self.on_gpu = False
        self.hyp = hyp
        self.qst_size = qst_size
        self.in_size = in_size
        self.out_size = out_size

    def cuda(self, device=None):
        self.on_gpu = True
        return super().cuda(device)


class RelationalLayer(RelationalLayerBase):
    """g over all ordered pairs, pair-sum, f (reference model.py:81-162).

    ``precision``: 'parity' (default where the tcgen05 kernels support the shape), 'fast', or 'fp32'
    (SIMT kernels, any shape).  Override globally with RN_B200_PRECISION."""

    def __init__(self, in_size, out_size, qst_size, hyp, extraction=False):
        super().__init__(in_size, out_size, qst_size, hyp)
        self.quest_inject_position = hyp["question_injection_position"]
        self.in_size = in_size
        self.g_layers_size = hyp["g_layers"]
        layers = []
        for idx, width in enumerate(hyp["g_layers"]):
            in_s = in_size if idx == 0 else hyp["g_layers"][idx - 1]
            if idx == self.quest_inject_position:
                in_s += qst_size
            layers.append(nn.Linear(in_s, width))
        self.g_layers = nn.ModuleList(layers)
        self.extraction = extraction
        self.precision = os.environ.get("RN_B200_PRECISION", "auto")
        self.dropout_mask_override = None     # tests: uint8 [B, f_fc2] keep-mask used instead of a fresh draw

    def _resolve_precision(self, n: int, k: int) -> str:
        if self.precision != "auto":
            return self.precision
        widths = self.g_layers_size
        uniform = all(w == widths[0] for w in widths)
        ok = uniform and ops.tc_supported(n, widths[0], len(widths), k, self.qst_size, self.quest_inject_position)
        return "parity" if ok else "fp32"

    def relation(self, x: torch.Tensor, qst: torch.Tensor) -> torch.Tensor:
        """x_g [B,G]: the fused pair construction + g-MLP + pair-sum."""
        if len(set(self.g_layers_size)) != 1:
            raise RuntimeError("the CUDA relation op needs equal-width g layers (all reference configs have them)")
        wb = []
        for layer in self.g_layers:
            wb += [layer.weight, layer.bias]
        b, n, k = x.shape
        return ops.RelationFunction.apply(x, qst, self.quest_inject_position, self._resolve_precision(n, k), *wb)

    def _hooked(self) -> bool:
        return any(len(layer._forward_hooks) or len(layer._forward_pre_hooks) for layer in self.g_layers)

    def _forward_materialised(self, x: torch.Tensor, qst: torch.Tensor) -> torch.Tensor:
        """Feature-extraction path (reference extract.py:40-47 hooks ``rl.g_layers[i]``): the fp32 kernels keep every
        g-layer activation in device memory and each hooked layer's hooks are called with the tensor the reference's
        nn.Linear would have received -- [B*n*n, fan_in] including the question columns at the injection layer.  The
        hook's `output` argument is the layer's post-ReLU activation (the pre-activation is never stored)."""
        wb = []
        for layer in self.g_layers:
            wb += [layer.weight, layer.bias]
        x_g, acts = ops.relation_forward_materialised(x, qst, self.quest_inject_position, *wb)
        b, n, k = x.shape
        for idx, layer in enumerate(self.g_layers):
            if not (len(layer._forward_hooks) or len(layer._forward_pre_hooks)):
                continue
            if idx == 0:      # the literal pair matrix (model.py:112-127): row a*n + c = [x_c | x_a]
                inp = torch.cat([x.unsqueeze(1).expand(b, n, n, k), x.unsqueeze(2).expand(b, n, n, k)], 3).reshape(b * n * n, 2 * k)
            else:
                inp = acts[idx - 1]
            if idx == self.quest_inject_position:
                inp = torch.cat([inp, qst.unsqueeze(1).expand(b, n * n, qst.shape[1]).reshape(b * n * n, -1)], 1)
            for hook in layer._forward_pre_hooks.values():
                hook(layer, (inp,))
            for hook in layer._forward_hooks.values():
                hook(layer, (inp,), acts[idx])
        return x_g

    def extract(self, x: torch.Tensor, qst: torch.Tensor, layer_idx: int):
        """(maxf, avgf) [B, width]: the aggregation of reference extract.py:63-74 for the input of g layer `layer_idx` --
        per sample the max and the mean over all pair rows of the L2-normalised rows (question columns stripped at the
        injection layer) -- computed by one reduction kernel over the materialised activation."""
        got = {}
        handle = self.g_layers[layer_idx].register_forward_hook(lambda m, i, o: got.__setitem__("z", i[0]))
        try:
            self._forward_materialised(x, qst)
        finally:
            handle.remove()
        z = got["z"]
        width = z.shape[1] - (self.qst_size if layer_idx == self.quest_inject_position else 0)
        return ops.extract_stats(z, x.shape[0], width)

    def forward(self, x: torch.Tensor, qst: torch.Tensor):
        if self.extraction or self._hooked():
            x_g = self._forward_materialised(x, qst)
            if self.extraction:
                return None              # like the reference (model.py:147-148): extraction stops after g
        else:
            x_g = self.relation(x, qst)
        p = self.dropout.p
        mask = None
        if self.training and p > 0 and self.dropout_mask_override is not None:
            mask = self.dropout_mask_override.to(device=x_g.device, dtype=torch.uint8).contiguous()
        elif self.training and p > 0:
            # keep-mask ~ Bernoulli(1 - p) from torch's CUDA generator (philox: CUDA-graph safe), one launch
            mask = torch.empty(x_g.shape[0], self.f_fc2.out_features, dtype=torch.uint8, device=x_g.device).bernoulli_(1.0 - p)
        return ops.FHeadFunction.apply(x_g, self.f_fc1.weight, self.f_fc1.bias, self.f_fc2.weight, self.f_fc2.bias,
                                       self.f_fc3.weight, self.f_fc3.bias, mask, 1.0 / (1.0 - p) if p < 1 else 0.0)


class RN(nn.Module):
    """The Relation Network (reference model.py:164-223): ``RN(args, hyp, extraction=False)`` with
    ``args.qdict_size`` / ``args.adict_size`` and the ``hyp`` dict of config.json;
    ``forward(img, qst_idxs)`` -> [B, adict] log-probabilities."""

    def __init__(self, args, hyp, extraction=False):
        super().__init__()
        self.coord_tensor = None      # kept for surface compatibility; coords are generated in-kernel
        self.on_gpu = False
        self.conv = ConvInputModel()
        self.state_desc = hyp["state_description"]
        hidden_size = hyp["lstm_hidden"]
        self.text = QuestionEmbedModel(args.qdict_size, embed=hyp["lstm_word_emb"], hidden=hidden_size)
        self.rl_in_size = hyp["rl_in_size"]
        self.rl_out_size = args.adict_size
        self.rl = RelationalLayer(self.rl_in_size, self.rl_out_size, hidden_size, hyp, extraction)
        if hyp["question_injection_position"] != 0:
            print("Supposing IR model")
        else:
            print("Supposing original DeepMind model")

    def forward(self, img: torch.Tensor, qst_idxs: torch.Tensor):
        if not img.is_cuda:
            raise RuntimeError("RN (B200-native) needs CUDA inputs: call model.cuda() and move the batch to the GPU")
        if self.state_desc or os.environ.get("RN_B200_TEXT_STREAM", "1") == "0" or (self.text.use_kernel and self.text.hidden == 128):
            # (the question-encoder kernel is one short launch: nothing to gain from a side stream, and the step stays
            # a single-stream sequence that CUDA-graph capture takes as is)
            ops.fork_point(img.device)             # the question encoder only needs the tokens: it may start now ...
            with ops.defer_text_join():
                qst = self.text(qst_idxs)          # ... on the auxiliary stream (ops.QuestionEncoderFunction), small batches
            x = img if self.state_desc else self._objects(img)
            ops.join_aux(img.device)               # the relation op reads q
            return self.rl(x, qst)
        # The question encoder (PyTorch/cuDNN, a chain of small latency-bound kernels) runs on a side stream next to
        # the conv stack; autograd replays its backward on the same side stream, next to the conv backward.
        cur = torch.cuda.current_stream(img.device)
        side = ops.side_stream(img.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            qst = self.text(qst_idxs)
        x = self._objects(img)
        cur.wait_stream(side)
        qst.record_stream(cur)
        return self.rl(x, qst)

    def _objects(self, img: torch.Tensor) -> torch.Tensor:
        if len(self.conv._forward_hooks) or len(self.conv._forward_pre_hooks):
            self.conv(img)                    # through the module: hooks on `conv` see the [B,24,d,d] feature map (extract.py:47)
            return self.conv.last_objects
        return self.conv.objects(img)

    def build_coord_tensor(self, b, d):
        """Reference-compatible helper (model.py:208-218); the kernels do not use it."""
        coords = torch.linspace(-d / 2.0, d / 2.0, d)
        x = coords.unsqueeze(0).repeat(d, 1)
        y = coords.unsqueeze(1).repeat(1, d)
        ct = torch.stack((x, y)).unsqueeze(0).repeat(b, 1, 1, 1)
        self.coord_tensor = ct.cuda() if self.on_gpu else ct

    def cuda(self, device=None):
        self.on_gpu = True
        self.rl.on_gpu = True
        return super().cuda(device)
