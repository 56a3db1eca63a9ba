"""Step-level host logic on the GPU: the gradient sink (gradients written in place into the optimiser's flat buffer) and
the CUDA-graphed training step must give the same parameters as the plain eager step."""
import contextlib
import io

import pytest
import torch

import relationnetworks_clevr_b200 as R
from oracle import rn_oracle as O
from relationnetworks_clevr_b200 import _lib, ops
from relationnetworks_clevr_b200.trainer import FlatClipAdam, GraphedTrainStep, train_step

pytestmark = pytest.mark.gpu
DEV = "cuda"


class _Args:
    qdict_size, adict_size = 82, 28


def _model(seed=5):
    hyp = O.HYPERPARAMS["original-fp"]
    with contextlib.redirect_stdout(io.StringIO()):
        m = R.RN(_Args, hyp)
    m.load_state_dict(O.seeded_params(hyp, 82, 28, seed), strict=False)
    m.to(DEV).train()
    B = 6
    m.rl.dropout_mask_override = (torch.rand(B, hyp["f_fc2"], generator=torch.Generator().manual_seed(1)) > 0.5).to(torch.uint8).to(DEV)
    batches = [(O.uniform_images(B, 128, 10 + i).to(DEV), O.questions(B, 20, 82, 20 + i).to(DEV), O.labels(B, 28, 30 + i).to(DEV))
               for i in range(3)]
    return m, batches


def _run(mode):
    ops.clear_grad_sink()
    m, batches = _model()
    opt = FlatClipAdam(m.parameters(), lr=1e-3, sink=(mode != "plain"))
    losses = []
    if mode == "graph":
        g = GraphedTrainStep(m, opt, *batches[0])
        assert g.captured, "CUDA-graph capture of the training step failed"
        before = _lib.lib().rn_launch_count()
        for b in batches:
            losses.append(float(g.step(*b)))
        assert _lib.lib().rn_launch_count() == before      # replays launch nothing from the host side of the library
    else:
        for b in batches:
            losses.append(float(train_step(m, opt, *b)))
    assert opt.step_count == len(batches)
    out = opt.flat.detach().clone(), [b.detach().clone() for b in m.buffers()], losses
    ops.clear_grad_sink()
    return out


def test_sink_and_graph_match_plain_step():
    flat_p, bufs_p, loss_p = _run("plain")
    flat_s, bufs_s, loss_s = _run("sink")
    flat_g, bufs_g, loss_g = _run("graph")
    assert loss_p == loss_s == loss_g
    assert torch.equal(flat_p, flat_s)          # same kernels, same order: bitwise
    assert torch.equal(flat_p, flat_g)
    for a, b, c in zip(bufs_p, bufs_s, bufs_g):
        assert torch.equal(a, b) and torch.equal(a, c)


def test_sink_leaves_param_grad_empty_and_covers_every_parameter():
    ops.clear_grad_sink()
    m, batches = _model()
    opt = FlatClipAdam(m.parameters(), lr=1e-3)
    opt.grad.fill_(float("nan"))
    opt.zero_grad()
    loss = torch.nn.functional.nll_loss(m(*batches[0][:2]), batches[0][2])
    loss.backward()
    assert all(p.grad is None for p in m.parameters())
    assert not torch.isnan(opt.grad).any()      # every slice was overwritten by exactly one kernel
    opt.check_aliasing()
    ops.clear_grad_sink()
