"""Diagnostics (not a test): a few full training steps (argv: steps, config, batch) for `ncu --metrics gpu__time_duration.sum`
launch lists; profiles/summarize.py `laststep` keeps the launches of the last step."""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import relationnetworks_clevr_b200 as R
from relationnetworks_clevr_b200.trainer import FlatClipAdam, train_step

class A: qdict_size, adict_size = 82, 28
model_name = sys.argv[2] if len(sys.argv) > 2 else "original-fp"
hyp = json.load(open(os.path.join(os.path.dirname(__file__), "..", "config.json")))["hyperparams"][model_name]
torch.manual_seed(42)
with contextlib.redirect_stdout(io.StringIO()):
    m = R.RN(A, hyp)
m.cuda().train()
opt = FlatClipAdam(m.parameters())
B = int(sys.argv[3]) if len(sys.argv) > 3 else 640
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
img = torch.rand(B, 3, 128, 128, device="cuda"); q = torch.randint(1, 83, (B, 20), device="cuda"); lab = torch.randint(0, 28, (B,), device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(steps):
    if i == steps - 1: e0.record()
    loss = train_step(m, opt, img, q, lab)
e1.record(); torch.cuda.synchronize()
print(f"{model_name}: last step {e0.elapsed_time(e1):.3f} ms, loss {float(loss):.4f}")
