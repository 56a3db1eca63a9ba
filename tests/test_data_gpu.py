"""PinnedBatchStager: every staged batch arrives bit-exact and in order while the next one is in flight."""
import pytest
import torch

from relationnetworks_clevr_b200 import data as D

pytestmark = pytest.mark.gpu


def test_stager_delivers_batches_in_order():
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(0)
    batches = [(torch.rand(16, 3, 32, 32, generator=g), torch.randint(1, 83, (16, 20), generator=g),
                torch.randint(0, 28, (16,), generator=g)) for _ in range(5)]
    stager = D.PinnedBatchStager(dev)
    seen = 0
    for i, (img, qst, lab) in enumerate(stager.iterate(batches)):
        assert img.is_cuda and qst.is_cuda and lab.is_cuda
        # consume on the compute stream (clone before the slot is recycled)
        got = (img.clone(), qst.clone(), lab.clone())
        torch.cuda.synchronize()
        for a, b in zip(got, batches[i]):
            assert torch.equal(a.cpu(), b)
        seen += 1
    assert seen == len(batches)
    assert list(stager.iterate([])) == []
