"""Generates tests/golden/data_contract.npz by running the UNMODIFIED reference utils.load_tensor_data
(/root/reference/utils.py:133-150) on a small seeded batch.  Run in the build container only."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
import utils as ref_utils  # noqa: E402

g = torch.Generator().manual_seed(7)
B, T = 5, 11
question = torch.zeros(B, T, dtype=torch.int64)
for i in range(B):
    n = int(torch.randint(3, T + 1, (1,), generator=g))
    question[i, :n] = torch.randint(1, 83, (n,), generator=g)
batch = {"image": torch.rand(B, 3, 8, 8, generator=g), "question": question,
         "answer": torch.randint(1, 29, (B, 1), generator=g, dtype=torch.int64)}
out = {k: v.numpy() for k, v in batch.items()}
for inv in (True, False):
    img, qst, label = ref_utils.load_tensor_data(batch, False, inv)
    tag = "inv" if inv else "fwd"
    out[f"img_{tag}"], out[f"qst_{tag}"], out[f"label_{tag}"] = img.numpy(), qst.numpy(), label.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_contract.npz"), **out)
print("wrote data_contract.npz", {k: v.shape for k, v in out.items()})
