"""Two-rank NCCL step == single-GPU step on the same global batch (needs >= 2 GPUs; skipped otherwise).
BatchNorm statistics are per rank (DataParallel semantics), so the single-GPU comparison runs the two shards
through the model separately and averages the gradients."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import rn_oracle as O

pytestmark = pytest.mark.gpu
B = 8


class _Args:
    qdict_size, adict_size = 82, 28


def _model(dev):
    import relationnetworks_clevr_b200 as R
    hyp = O.HYPERPARAMS["original-fp"]
    m = R.RN(_Args, hyp)
    m.load_state_dict(O.seeded_params(hyp, 82, 28, seed=9), strict=False)
    m.to(dev).train()
    m.rl.precision = "fp32"                      # exact arithmetic: the test is about the exchange step
    m.rl.dropout_mask_override = torch.ones(B // 2, hyp["f_fc2"], dtype=torch.uint8, device=dev)      # (on the device: capturable)
    return m


def _batch():
    return O.uniform_images(B, 128, seed=1), O.questions(B, 9, 82, seed=2), O.labels(B, 28, seed=3)


def _worker(rank, world, init_file, out_file, graph):
    from relationnetworks_clevr_b200.trainer import FlatClipAdam, GraphedTrainStep, shard_rows, train_step
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world, device_id=dev)
    m = _model(dev)
    opt = FlatClipAdam(m.parameters(), lr=1e-3)
    img, qst, lab = _batch()
    rows = shard_rows(B, rank, world)
    batch = (img[rows].to(dev), qst[rows].to(dev), lab[rows].to(dev))
    if graph:              # the whole step, NCCL all-reduce included, captured once and replayed
        g = GraphedTrainStep(m, opt, *batch)
        assert g.captured, "CUDA-graph capture (with the NCCL all-reduce inside) failed"
        g.step(*batch)
    else:
        train_step(m, opt, *batch)
    torch.cuda.synchronize()
    assert opt.step_count == 1
    torch.save(opt.flat.cpu(), f"{out_file}.{rank}")
    if graph:
        del g                  # the captured graph holds NCCL work: release it before the communicator goes away
    dist.barrier()
    torch.cuda.synchronize()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("graph", [False, True])
def test_two_rank_nccl_step_matches_single_gpu(graph):
    from relationnetworks_clevr_b200.trainer import FlatClipAdam, shard_rows
    from relationnetworks_clevr_b200 import ops
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file, out_file = os.path.join(tmp, "init"), os.path.join(tmp, "out")
        mp.spawn(_worker, args=(world, init_file, out_file, graph), nprocs=world, join=True)
        got = [torch.load(f"{out_file}.{r}") for r in range(world)]
    assert torch.equal(got[0], got[1])
    dev = torch.device("cuda", 0)
    m = _model(dev)
    opt = FlatClipAdam(m.parameters(), lr=1e-3)
    img, qst, lab = _batch()
    total = torch.zeros_like(opt.grad)
    for r in range(world):
        rows = shard_rows(B, r, world)
        opt.zero_grad()
        F.nll_loss(m(img[rows].to(dev), qst[rows].to(dev)), lab[rows].to(dev)).backward()
        total += opt.gather_grads() / world
    ops.clip_adam_(opt.flat, total, opt.exp_avg, opt.exp_avg_sq, 1, 1e-3)
    assert O.rel_err(got[0], opt.flat.cpu()) < 1e-5
    ops.clear_grad_sink()
