#!/usr/bin/env python
"""Thin driver keeping the reference's `train.py` command-line surface on top of the B200-native model.

Same flags and defaults as the reference (train.py:369-415), same `--model original-fp|original-sd|ir-fp|ir-sd`
selection from `config.json`, same run-directory / checkpoint naming (`RN_epoch_NN.pth`, train.py:363-364), same
log line formats (train.py:61-63, 138-145 -- `plot.py`-scrapable).  The step body is the reference's
(train.py:39-48): zero_grad, forward, nll_loss, backward, clip_grad_norm, Adam -- executed by
`relationnetworks_clevr_b200` (one process per GPU under torchrun replaces `nn.DataParallel`, train.py:256-258).

The CLEVR dataset connectors (`clevr_dataset_connector.py`, `utils.py`) are outside the hot path
and there is no dataset in this environment: `--synthetic-batches N` (default 20) feeds CLEVR-shaped synthetic
batches through the same tensor contract as `utils.load_tensor_data` (utils.py:133-150).
"""
from __future__ import annotations

import argparse
import json
import os
import re

import torch
import torch.distributed as dist
import torch.nn.functional as F

from relationnetworks_clevr_b200 import RN
from relationnetworks_clevr_b200.trainer import FlatClipAdam, train_step

QDICT_SIZE, ADICT_SIZE = 82, 28          # CLEVR vocabulary sizes pinned by the shipped checkpoints


def synthetic_batches(args, hyp, n_batches, batch_size, device, seed):
    """CLEVR-shaped batches: images [B,3,128,128] in [0,1) (or [B,12,7] state descriptions), reversed questions
    left-padded with 0 (utils.py:138-141), labels already shifted to 0..27 (utils.py:149)."""
    g = torch.Generator().manual_seed(seed)
    for _ in range(n_batches):
        if hyp["state_description"]:
            img = torch.zeros(batch_size, 12, 7)
            img[:, :10] = torch.rand(batch_size, 10, 7, generator=g) * 3
        else:
            img = torch.rand(batch_size, 3, 128, 128, generator=g)
        qst = torch.randint(1, QDICT_SIZE + 1, (batch_size, 20), generator=g)
        if args.invert_questions:
            qst[:, : int(torch.randint(0, 8, (1,), generator=g))] = 0
        label = torch.randint(0, ADICT_SIZE, (batch_size,), generator=g)
        yield img.to(device), qst.to(device), label.to(device)


def train(batches, n_batches, model, optimizer, epoch, args):
    model.train()
    avg_loss, n_seen = 0.0, 0
    for batch_idx, (img, qst, label) in enumerate(batches):
        loss = train_step(model, optimizer, img, qst, label)
        avg_loss += float(loss.detach())
        n_seen += 1
        if batch_idx % args.log_interval == 0:
            processed = batch_idx * args.batch_size
            n_samples = n_batches * args.batch_size
            print('Train Epoch: {} [{}/{} ({:.0%})] Train loss: {}'.format(
                epoch, processed, n_samples, float(processed) / n_samples, avg_loss / n_seen))
            avg_loss, n_seen = 0.0, 0


def test(batches, model, epoch):
    model.eval()
    corrects, n_samples, loss_sum, n_batches = 0, 0, 0.0, 0
    with torch.no_grad():
        for img, qst, label in batches:
            output = model(img, qst)
            loss_sum += float(F.nll_loss(output, label))
            corrects += int((output.argmax(1) == label).sum())
            n_samples += label.numel()
            n_batches += 1
    accuracy = corrects / max(n_samples, 1)
    print('Test Epoch {}: Accuracy = {:.2%} ({:g}/{}); Test loss = {}'.format(
        epoch, accuracy, corrects, n_samples, loss_sum / max(n_batches, 1)))
    return accuracy


def main(args):
    with open(args.config) as config_file:
        hyp = json.load(config_file)['hyperparams'][args.model]
    if args.dropout > 0:
        hyp['dropout'] = args.dropout
    if args.question_injection >= 0:
        hyp['question_injection_position'] = args.question_injection
    print('Loaded hyperparameters from configuration {}, model: {}: {}'.format(args.config, args.model, hyp))

    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or args.no_cuda:
        raise SystemExit("this driver runs the sm_100a kernels: a CUDA device is required (no CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    rank = dist.get_rank() if world > 1 else 0

    args.model_dirs = ('./model_{}_drop{}_bstart{}_bstep{}_bgamma{}_bmax{}_lrstart{}_'
                       'lrstep{}_lrgamma{}_lrmax{}_invquests-{}_clipnorm{}_glayers{}_qinj{}_fc1{}_fc2{}').format(
        args.model, hyp['dropout'], args.batch_size, args.bs_step, args.bs_gamma, args.bs_max, args.lr, args.lr_step,
        args.lr_gamma, args.lr_max, args.invert_questions, args.clip_norm, hyp['g_layers'],
        hyp['question_injection_position'], hyp['f_fc1'], hyp['f_fc2'])
    if rank == 0:
        os.makedirs(args.model_dirs, exist_ok=True)
        with open(os.path.join(args.model_dirs, 'config.txt'), 'w') as f:
            f.write(str(args) + '\n\n' + str(hyp))

    torch.manual_seed(args.seed)
    torch.cuda.manual_seed(args.seed)
    args.qdict_size, args.adict_size = QDICT_SIZE, ADICT_SIZE
    model = RN(args, hyp)
    model.cuda()

    start_epoch = 1
    if args.resume:
        checkpoint = torch.load(args.resume, map_location="cpu", weights_only=True)
        checkpoint = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in checkpoint.items()}
        model.load_state_dict(checkpoint, strict=True)         # like the reference: a key mismatch is an error
        m = re.search(r'epoch_(\d+)', os.path.basename(args.resume))
        start_epoch = int(m.group(1)) + 1 if m else 1
        print('==> loaded checkpoint {} (next epoch {})'.format(args.resume, start_epoch))
    if args.conv_transfer_learn:
        checkpoint = torch.load(args.conv_transfer_learn, map_location="cpu", weights_only=True)
        conv = {k.split("conv.", 1)[1]: v for k, v in checkpoint.items() if "conv." in k}
        model.conv.load_state_dict(conv, strict=True)
        print('==> loaded conv layers from {}'.format(args.conv_transfer_learn))

    if args.batch_size % world:
        raise SystemExit('--batch-size {} is not divisible by the number of ranks {}'.format(args.batch_size, world))
    per_rank_bs = args.batch_size // world
    # identical parameters on every rank (seeded above), but independent dropout masks per rank
    torch.cuda.manual_seed(args.seed + rank)
    if args.test:
        test(synthetic_batches(args, hyp, args.synthetic_batches, per_rank_bs, device, args.seed + 1), model, start_epoch)
        return

    lr = args.lr * (args.lr_gamma ** ((start_epoch - 1) // args.lr_step))
    optimizer = FlatClipAdam(model.parameters(), lr=min(lr, args.lr_max), weight_decay=1e-4, clip_norm=args.clip_norm)
    print('Training ({} epochs) is starting...'.format(args.epochs))
    bs = args.batch_size
    for epoch in range(start_epoch, args.epochs + 1):
        # batch-size schedule of the reference (train.py:337-341): grows by bs_gamma every bs_step epochs up to bs_max
        if ((args.bs_max > 0 and bs < args.bs_max) or args.bs_max < 0) and (epoch % args.bs_step == 0 or epoch == start_epoch):
            bs = int(args.batch_size * (args.bs_gamma ** (epoch // args.bs_step)))
            if args.bs_max > 0:
                bs = min(bs, args.bs_max)
            bs -= bs % world
            per_rank_bs = bs // world
            print('Dataset reinitialized with batch size {}'.format(bs))
        # StepLR-like schedule: the LR is multiplied by lr_gamma every lr_step epochs WHILE it is below lr_max
        # (train.py:332-333, 350-351: the reference steps its scheduler only while get_lr() < lr_max)
        if (epoch - 1) % args.lr_step == 0 and epoch > 1 and (args.lr_max < 0 or optimizer.lr < args.lr_max):
            optimizer.set_lr(optimizer.lr * args.lr_gamma)
        print('Current learning rate: {}'.format(optimizer.lr))
        train(synthetic_batches(args, hyp, args.synthetic_batches, per_rank_bs, device, args.seed + 100 * epoch + rank),
              args.synthetic_batches, model, optimizer, epoch, args)
        test(synthetic_batches(args, hyp, max(1, args.synthetic_batches // 4), per_rank_bs, device, args.seed + 7), model, epoch)
        if rank == 0:
            torch.save(model.state_dict(), os.path.join(args.model_dirs, 'RN_epoch_{:02d}.pth'.format(epoch)))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='B200-native Relational-Network CLEVR (reference train.py surface)')
    parser.add_argument('--batch-size', type=int, default=640, metavar='N')
    parser.add_argument('--test-batch-size', type=int, default=640)
    parser.add_argument('--epochs', type=int, default=350, metavar='N')
    parser.add_argument('--lr', type=float, default=0.000005, metavar='LR')
    parser.add_argument('--clip-norm', type=int, default=50)
    parser.add_argument('--no-cuda', action='store_true', default=False)
    parser.add_argument('--seed', type=int, default=42, metavar='S')
    parser.add_argument('--log-interval', type=int, default=10, metavar='N')
    parser.add_argument('--resume', type=str)
    parser.add_argument('--clevr-dir', type=str, default='.')
    parser.add_argument('--model', type=str, default='original-fp')
    parser.add_argument('--no-invert-questions', action='store_true', default=False)
    parser.add_argument('--test', action='store_true', default=False)
    parser.add_argument('--conv-transfer-learn', type=str)
    parser.add_argument('--lr-max', type=float, default=0.0005)
    parser.add_argument('--lr-gamma', type=float, default=2)
    parser.add_argument('--lr-step', type=int, default=20)
    parser.add_argument('--bs-max', type=int, default=-1)
    parser.add_argument('--bs-gamma', type=float, default=1)
    parser.add_argument('--bs-step', type=int, default=20)
    parser.add_argument('--dropout', type=float, default=-1)
    parser.add_argument('--config', type=str, default='config.json')
    parser.add_argument('--question-injection', type=int, default=-1)
    parser.add_argument('--synthetic-batches', type=int, default=20,
                        help='synthetic CLEVR-shaped batches per epoch (no dataset connector in this build)')
    args = parser.parse_args()
    args.invert_questions = not args.no_invert_questions
    main(args)
