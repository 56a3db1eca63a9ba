"""Diagnostics (not a test): host-side enqueue time of one training step vs its GPU time."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import relationnetworks_clevr_b200 as R
from relationnetworks_clevr_b200.trainer import FlatClipAdam, train_step

class A: qdict_size, adict_size = 82, 28
hyp = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config.json")))["hyperparams"]["original-fp"]
torch.manual_seed(42)
with contextlib.redirect_stdout(io.StringIO()):
    m = R.RN(A, hyp)
m.cuda().train()
opt = FlatClipAdam(m.parameters())
B = 640
img = torch.rand(B, 3, 128, 128, device="cuda"); q = torch.randint(1, 83, (B, 20), device="cuda"); lab = torch.randint(0, 28, (B,), device="cuda")
for _ in range(5): train_step(m, opt, img, q, lab)
torch.cuda.synchronize()
ts = []
for _ in range(20):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); train_step(m, opt, img, q, lab); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    ts.append((t1 - t0, t2 - t0))
print("cpu enqueue ms (median, max):", sorted(t[0] for t in ts)[10] * 1e3, max(t[0] for t in ts) * 1e3)
print("step ms incl. sync (median):", sorted(t[1] for t in ts)[10] * 1e3)
