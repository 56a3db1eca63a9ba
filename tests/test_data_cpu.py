"""Host-side tensor contracts (reference utils.py:71-150): collate padding, question inversion, label shift."""
import numpy as np
import pytest
import torch

from relationnetworks_clevr_b200 import data as D


def _sample(n_tokens, seed, sd_objects=None):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(sd_objects, 7, generator=g) if sd_objects else torch.rand(3, 8, 8, generator=g)
    return {"image": img, "question": torch.randint(1, 83, (n_tokens,), generator=g, dtype=torch.int64),
            "answer": torch.randint(1, 29, (1,), generator=g, dtype=torch.int64)}


def test_collate_from_pixels_pads_questions_right_with_zero():
    batch = [_sample(5, 0), _sample(9, 1), _sample(3, 2)]
    out = D.collate_samples_from_pixels(batch)
    assert out["image"].shape == (3, 3, 8, 8) and out["answer"].shape == (3, 1)
    q = out["question"]
    assert q.shape == (3, 9) and q.dtype == torch.int64
    for i, s in enumerate(batch):
        n = len(s["question"])
        assert torch.equal(q[i, :n], s["question"]) and int(q[i, n:].abs().sum()) == 0


def test_collate_state_description_pads_objects_to_12():
    batch = [_sample(4, 3, sd_objects=10), _sample(4, 4, sd_objects=3)]
    out = D.collate_samples_state_description(batch)
    assert out["image"].shape == (2, 12, 7)
    assert torch.equal(out["image"][1, :3], batch[1]["image"]) and float(out["image"][1, 3:].abs().sum()) == 0.0
    only = D.collate_samples_images_state_description([b["image"] for b in batch])
    assert torch.equal(only, out["image"])


def test_load_tensor_data_inverts_and_shifts_labels():
    batch = D.collate_samples_from_pixels([_sample(5, 0), _sample(9, 1)])
    img, qst, label = D.load_tensor_data(batch, cuda=False, invert_questions=True)
    assert torch.equal(qst, batch["question"].flip(1))
    assert int(qst[0, :4].abs().sum()) == 0 and int(qst[0, 4]) != 0          # right padding became LEFT padding
    assert label.shape == (2,) and torch.equal(label, batch["answer"].squeeze(1) - 1)
    assert img is batch["image"]
    _, q2, _ = D.load_tensor_data(batch, cuda=False, invert_questions=False, volatile=True)
    assert torch.equal(q2, batch["question"])


def test_load_tensor_data_matches_reference_golden():
    """tests/golden/data_contract.npz was produced by the reference's own utils.load_tensor_data
    (tests/golden/make_golden_data.py)."""
    z = np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "data_contract.npz"))
    batch = {"image": torch.from_numpy(z["image"]), "question": torch.from_numpy(z["question"]),
             "answer": torch.from_numpy(z["answer"])}
    for inv in (True, False):
        img, qst, label = D.load_tensor_data(batch, cuda=False, invert_questions=inv)
        tag = "inv" if inv else "fwd"
        assert torch.equal(qst, torch.from_numpy(z[f"qst_{tag}"]))
        assert torch.equal(label, torch.from_numpy(z[f"label_{tag}"]))
        assert torch.equal(img, torch.from_numpy(z[f"img_{tag}"]))


def test_stager_needs_cuda():
    with pytest.raises(RuntimeError):
        D.PinnedBatchStager(torch.device("cpu"))
