"""Diagnostics (not a test): a few eager training steps at a given batch, for `ncu` launch lists.
    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <n> --csv --log-file out.csv python tests/diag_step_profile.py 80 3
"""
import contextlib
import io
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import relationnetworks_clevr_b200 as R
from relationnetworks_clevr_b200.trainer import FlatClipAdam, train_step

B = int(sys.argv[1]) if len(sys.argv) > 1 else 640
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
config = sys.argv[3] if len(sys.argv) > 3 else "original-fp"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
hyp = json.load(open(os.path.join(root, "config.json")))["hyperparams"][config]


class A:
    qdict_size, adict_size = 82, 28


torch.manual_seed(0)
with contextlib.redirect_stdout(io.StringIO()):
    m = R.RN(A, hyp)
m.cuda().train()
opt = FlatClipAdam(m.parameters())
g = torch.Generator().manual_seed(1)
img = torch.rand(B, 3, 128, 128, generator=g).cuda()
qst = torch.randint(1, 83, (B, 20), generator=g).cuda()
lab = torch.randint(0, 28, (B,), generator=g).cuda()
for _ in range(steps):
    train_step(m, opt, img, qst, lab)
torch.cuda.synchronize()
print("done", B, steps)
