"""Host-side tensor contracts either side of the hot path (reference utils.py:71-120 collate, :133-150
load_tensor_data), restated for current PyTorch, plus the pinned double-buffered host->device stager the
end-to-end benchmark loop uses.

The reference's own ``collate_samples`` ends in ``torch.stack(padded_questions)`` on a 2-D tensor, which PyTorch
0.3.1 accepted (stacking the rows back into the same [B, T] tensor) and PyTorch >= 1.0 rejects; the contract kept
here is the 0.3.1 behaviour: ``question`` is the zero-padded [B, max_len] int64 tensor itself.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

SD_MAX_OBJECTS = 12        # utils.py:101: state descriptions are padded to 12 objects


def collate_samples(batch: Sequence, state_description: bool, only_images: bool):
    """Merge samples into one mini-batch (utils.py:78-117).

    Samples are dicts ``{'image', 'question', 'answer'}`` (or bare images when ``only_images``).  Questions are
    right-padded with index 0 to the longest question of the batch; state-description object matrices [n_i, F] are
    zero-padded to [12, F]."""
    batch_size = len(batch)
    if only_images:
        images = list(batch)
    else:
        images = [d["image"] for d in batch]
        answers = [d["answer"] for d in batch]
        questions = [d["question"] for d in batch]
        max_len = max(len(q) for q in questions)
        padded_questions = torch.zeros(batch_size, max_len, dtype=torch.int64)
        for i, q in enumerate(questions):
            padded_questions[i, :len(q)] = q
    if state_description:
        feat = images[0].size(1)
        padded_objects = torch.zeros(batch_size, SD_MAX_OBJECTS, feat, dtype=torch.float32)
        for i, o in enumerate(images):
            padded_objects[i, :o.size(0), :] = o
        images = padded_objects
    if only_images:
        return images if torch.is_tensor(images) else torch.stack(images)
    return dict(image=images if torch.is_tensor(images) else torch.stack(images), answer=torch.stack(answers),
                question=padded_questions)


def collate_samples_from_pixels(batch):
    return collate_samples(batch, False, False)


def collate_samples_state_description(batch):
    return collate_samples(batch, True, False)


def collate_samples_images_state_description(batch):
    return collate_samples(batch, True, True)


def load_tensor_data(data_batch: Dict[str, torch.Tensor], cuda: bool, invert_questions: bool, volatile: bool = False,
                     device: Optional[torch.device] = None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(img, qst, label) as the training / test loops consume them (utils.py:133-150).

    ``invert_questions`` reverses the token order, which turns the collate's right padding into LEFT padding (the
    LSTM then ends on the first word); labels arrive as [B, 1] in 1..A and leave as [B] in 0..A-1.  ``volatile`` is
    accepted for signature compatibility (callers wrap evaluation in ``torch.no_grad()``)."""
    del volatile
    qst = data_batch["question"]
    if invert_questions:
        qst_len = qst.size(1)
        qst = qst.index_select(1, torch.arange(qst_len - 1, -1, -1, dtype=torch.int64))
    img, label = data_batch["image"], data_batch["answer"]
    if cuda:
        dev = device if device is not None else torch.device("cuda")
        img, qst, label = img.to(dev, non_blocking=True), qst.to(dev, non_blocking=True), label.to(dev, non_blocking=True)
    label = (label - 1).squeeze(1)
    return img, qst, label


class PinnedBatchStager:
    """Double-buffered host->device staging of (img, qst, label) batches: batch i+1 is copied from pinned host memory
    on a side stream while batch i is being consumed on the compute stream.

        stager = PinnedBatchStager(device)
        for img, qst, label in stager.iterate(batches):      # batches: iterable of CPU tensor triples
            loss = train_step(model, opt, img, qst, label)
    """

    def __init__(self, device: torch.device):
        if torch.device(device).type != "cuda":
            raise RuntimeError("PinnedBatchStager stages into CUDA memory")
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self._pinned: List[Optional[Tuple[torch.Tensor, ...]]] = [None, None]
        self._dev: List[Optional[Tuple[torch.Tensor, ...]]] = [None, None]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def _buffers(self, slot: int, batch: Tuple[torch.Tensor, ...]):
        cur = self._pinned[slot]
        if cur is None or any(c.shape != t.shape or c.dtype != t.dtype for c, t in zip(cur, batch)):
            self._pinned[slot] = tuple(torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in batch)
            self._dev[slot] = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch)
            self._consumed[slot].record(torch.cuda.current_stream(self.device))
        return self._pinned[slot], self._dev[slot]

    def _stage(self, slot: int, batch: Tuple[torch.Tensor, ...]) -> None:
        pinned, dev = self._buffers(slot, batch)
        self._consumed[slot].synchronize()          # the pinned buffer of this slot is about to be rewritten on the host
        for p, t in zip(pinned, batch):
            p.copy_(t)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self._consumed[slot])
            for d, p in zip(dev, pinned):
                d.copy_(p, non_blocking=True)
            self._ready[slot].record(self.copy_stream)

    def iterate(self, batches: Iterable[Tuple[torch.Tensor, ...]]) -> Iterator[Tuple[torch.Tensor, ...]]:
        it = iter(batches)
        try:
            nxt = next(it)
        except StopIteration:
            return
        self._stage(0, nxt)
        i = 0
        while True:
            slot = i % 2
            try:
                nxt = next(it)
                have_next = True
            except StopIteration:
                have_next = False
            if have_next:
                self._stage(1 - slot, nxt)
            torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
            yield self._dev[slot]
            self._consumed[slot].record(torch.cuda.current_stream(self.device))
            if not have_next:
                return
            i += 1
