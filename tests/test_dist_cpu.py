"""World-size-2 data-parallel protocol on CPU (gloo): shard -> local gradients (oracle) -> ONE flat all-reduce
-> scale 1/W -> clip + Adam must equal a single process stepping on the whole batch with per-shard BatchNorm
statistics (the reference's DataParallel semantics)."""
import os
import tempfile

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import rn_oracle as O
from relationnetworks_clevr_b200.trainer import allreduce_flat_, shard_rows

CONFIG = "original-fp"
SIDE, B = 32, 4          # 2x2 grid keeps the CPU oracle fast; BN statistics are per shard


def _inputs():
    img = O.uniform_images(B, SIDE, seed=5)
    qst = O.questions(B, 7, 82, seed=6)
    lab = O.labels(B, 28, seed=7)
    return img, qst, lab


def _flat_grads(params, names, img, qst, lab):
    leaves = {k: (v.clone().requires_grad_(True) if k in names else v.clone()) for k, v in params.items()}
    loss = F.nll_loss(O.rn_forward(leaves, O.HYPERPARAMS[CONFIG], img, qst, training=True,
                                   dropout_mask=torch.ones(img.shape[0], 256)), lab)
    loss.backward()
    return torch.cat([leaves[k].grad.reshape(-1) for k in names]), float(loss)


def _worker(rank, world, init_file, out_file):
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    torch.set_num_threads(2)
    hyp = O.HYPERPARAMS[CONFIG]
    params = O.seeded_params(hyp, 82, 28, seed=3)                 # identical init on every rank
    names = [k for k in params if "running" not in k]
    img, qst, lab = _inputs()
    rows = shard_rows(B, rank, world)
    flat, _ = _flat_grads(params, names, img[rows], qst[rows], lab[rows])
    scale = allreduce_flat_(flat)                                  # the single exchange step
    flat *= scale
    plist = [params[k].clone() for k in names]
    grads, off = [], 0
    for w in plist:
        grads.append(flat[off:off + w.numel()].view_as(w).clone())
        off += w.numel()
    m = [torch.zeros_like(w) for w in plist]
    v = [torch.zeros_like(w) for w in plist]
    O.clip_and_adam(plist, grads, m, v, 1, lr=1e-3)
    torch.save({"flat": flat, "params": torch.cat([w.reshape(-1) for w in plist])}, f"{out_file}.{rank}")
    dist.destroy_process_group()


def test_shard_rows():
    assert shard_rows(640, 3, 8) == slice(240, 320)
    with pytest.raises(ValueError):
        shard_rows(10, 0, 4)


def test_two_rank_step_equals_single_process_step():
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        init_file, out_file = os.path.join(tmp, "init"), os.path.join(tmp, "out")
        mp.spawn(_worker, args=(world, init_file, out_file), nprocs=world, join=True)
        res = [torch.load(f"{out_file}.{r}") for r in range(world)]
    # both ranks hold the same averaged gradient and the same updated parameters
    assert torch.equal(res[0]["flat"], res[1]["flat"])
    assert torch.equal(res[0]["params"], res[1]["params"])
    # single process: mean of the per-shard mean losses, BatchNorm statistics per shard
    hyp = O.HYPERPARAMS[CONFIG]
    params = O.seeded_params(hyp, 82, 28, seed=3)
    names = [k for k in params if "running" not in k]
    img, qst, lab = _inputs()
    ref = torch.zeros_like(res[0]["flat"])
    for r in range(world):
        rows = shard_rows(B, r, world)
        g, _ = _flat_grads(params, names, img[rows], qst[rows], lab[rows])
        ref += g / world
    assert O.rel_err(res[0]["flat"], ref) < 1e-6
    plist = [params[k].clone() for k in names]
    grads, off = [], 0
    for w in plist:
        grads.append(ref[off:off + w.numel()].view_as(w).clone())
        off += w.numel()
    O.clip_and_adam(plist, grads, [torch.zeros_like(w) for w in plist], [torch.zeros_like(w) for w in plist], 1, lr=1e-3)
    assert O.rel_err(res[0]["params"], torch.cat([w.reshape(-1) for w in plist])) < 1e-6
