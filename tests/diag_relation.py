"""Diagnostics (not a test): tcgen05 relation op vs the fp32 SIMT path on the same GPU, larger batches."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from relationnetworks_clevr_b200 import ops
from tests.test_parity_gpu import _g_params

def run(B, n, k, Q, G, qinj, precision, seed=0, scale=2.0):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(B, n, k, generator=gen); q = torch.randn(B, Q, generator=gen)
    gp = _g_params(n, k, Q, G, qinj, gen, scale=scale); dxg = torch.randn(B, G, generator=gen)
    xc, qc = x.cuda().requires_grad_(True), q.cuda().requires_grad_(True)
    wb = []
    for w, b in gp: wb += [w.cuda().requires_grad_(True), b.cuda().requires_grad_(True)]
    xg = ops.RelationFunction.apply(xc, qc, qinj, precision, *wb)
    xg.backward(dxg.cuda())
    out = {"xg": xg.detach(), "dx": xc.grad, "dq": qc.grad}
    for l in range(4): out[f"dW{l}"], out[f"db{l}"] = wb[2*l].grad, wb[2*l+1].grad
    return {k_: v.double().cpu() for k_, v in out.items()}

if __name__ == "__main__":
    for (B, n, qinj) in [(3, 16, 0), (32, 16, 0), (32, 64, 0), (32, 64, 2), (128, 64, 0)]:
        ref = run(B, n, 26, 128, 256, qinj, "fp32")
        for prec in sys.argv[1:] or ["parity"]:
            got = run(B, n, 26, 128, 256, qinj, prec)
            line = {}
            for k_ in ref:
                d = (got[k_] - ref[k_])
                line[k_] = "%.1e/%.1e" % (d.abs().max() / ref[k_].abs().max(), d.norm() / ref[k_].norm())
            print(f"B={B} n={n} qinj={qinj} {prec} (max/l2):", line, flush=True)
