#!/usr/bin/env python
"""Headline benchmark: questions/s of one Relation-Network TRAINING step (forward + backward +
gradient all-reduce + clip/Adam) on BASELINE.json config 2: original-fp, 128x128 images, 8x8x24 grid,
batch 640 per GPU, synthetic data, seeded random-init weights.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's algorithm (oracle port, PyTorch CPU) on host cores

Prints ONE JSON line (rank 0).  `value` = device-resident throughput; `e2e` = the same step through
the public API with pinned-host inputs copied every step and the loss read back; `roofline` = the g-MLP
(relation forward + backward launches) against the measured bf16 tensor peak; `cpu_baseline` = the
oracle port timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

QDICT, ADICT, T_Q = 82, 28, 20
CONFIG = "original-fp"
SIDE = 128
# algorithmic g-MLP FLOPs per sample (SURVEY.md 8d / BASELINE.md 2): 242,688 MAC/pair x 4096 pairs x 2 x 3 (fwd+bwd)
G_FLOP_FWD = 2.0 * 242_688 * 4096
G_FLOP_TRAIN = 3.0 * G_FLOP_FWD


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region, through NVML in a background thread (one cheap
    query every 50 ms).  An `nvidia-smi -lms` subprocess is only the fallback: its full-device queries take driver
    locks that stall kernel launches for tens of milliseconds each."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASON_BITS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.thread = index, [], None, None
        self._stop = threading.Event()
        self.nvml = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in vis.split(",") if v.strip().isdigit()]
            phys = int(ids[self.index]) if self.index < len(ids) else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "250"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((mhz, self.max_mhz, mask))
            except Exception:
                pass
            self._stop.wait(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            rows = list(self.rows)
            sm = [r[0] for r in rows]
            reasons = sorted({name for r in rows for bit, name in self.REASON_BITS if r[2] & bit})
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


def synthetic_batch(B: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, SIDE, SIDE, generator=g)                       # ToTensor range, no normalisation
    qst = torch.randint(1, QDICT + 1, (B, T_Q), generator=g, dtype=torch.int64)
    lab = torch.randint(0, ADICT, (B,), generator=g, dtype=torch.int64)
    return img, qst, lab


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's algorithm (oracle port, plain PyTorch CPU, materialised
# pairs + autograd) on the host cores.  The reference is Python and cannot travel to the GPU box, so this
# is kind="port" (oracle/rn_oracle.py, pinned to the reference by tests/golden).
# --------------------------------------------------------------------------------------------------
def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_port_qps(sample_B: int, steps: int, warmup: int, max_seconds: float = 30.0):
    from oracle import rn_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    hyp = O.HYPERPARAMS[CONFIG]
    p = O.seeded_params(hyp, QDICT, ADICT, seed=42)
    names = [k for k in p if "running" not in k]
    leaves = {k: (p[k].clone().requires_grad_(True) if k in names else p[k].clone()) for k in p}
    plist = [leaves[k] for k in names]
    m = [torch.zeros_like(w) for w in plist]
    v = [torch.zeros_like(w) for w in plist]
    img, qst, lab = synthetic_batch(sample_B, seed=42)
    times = []
    t_start = time.perf_counter()
    for it in range(warmup + steps):
        if len(times) >= 2 and time.perf_counter() - t_start > max_seconds:
            break                              # bounded sample: slow hosts stop early (at least two timed steps)
        t0 = time.perf_counter()
        for w in plist:
            w.grad = None
        loss = F.nll_loss(O.rn_forward(leaves, hyp, img, qst, training=True), lab)
        loss.backward()
        with torch.no_grad():
            O.clip_and_adam(plist, [w.grad for w in plist], m, v, it + 1, lr=5e-6)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    best = min(times)
    return sample_B / best, cores, best, len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample_B = 64
    steps = max(1, min(args.steps, 40))           # ~0.45 s per step on 16 host cores: the whole arm stays well under a minute
    warm = max(1, min(args.warmup, 5))
    qps, cores, best, steps = cpu_port_qps(sample_B, steps, warm, max_seconds=120.0)
    sample = f"{steps} timed full training steps (fwd+bwd+clip+Adam) at batch {sample_B} of the batch-640 workload, best step"
    line = {"impl": "reference", "metric": "questions/sec", "value": qps, "unit": "questions/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": best * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "original-fp training step, 128x128 images, 8x8 grid, 4096 pairs/sample, batch 640 (timed on a batch-64 sample)"},
            "cpu_baseline": {"value": qps, "unit": "questions/s", "cores": cores, "cpu": cpu_model_name(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": qps, "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import relationnetworks_clevr_b200 as R
    from relationnetworks_clevr_b200 import _lib, ops
    from relationnetworks_clevr_b200.trainer import FlatClipAdam, train_step

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: no NCCL version banner unless the caller asked for a verbose NCCL log
        if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE", "ABORT"):
            os.environ["NCCL_DEBUG"] = "NONE"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    _lib.check(_lib.lib().rn_device_check(local_rank), "rn_device_check")

    B = args.batch if args.scaling == "weak" else args.batch // world

    class A:
        qdict_size, adict_size = QDICT, ADICT

    hyp = json.load(open(os.path.join(ROOT, "config.json")))["hyperparams"][CONFIG]
    import contextlib
    import io
    torch.manual_seed(42)                       # train.py:382 -- identical random init on all ranks
    with contextlib.redirect_stdout(io.StringIO()):
        model = R.RN(A, hyp)
    model.to(dev).train()
    model.rl.precision = args.precision
    opt = FlatClipAdam(model.parameters(), lr=5e-6, weight_decay=1e-4, clip_norm=50.0)
    precision = model.rl._resolve_precision(64, 26)

    # several distinct device-resident batches (each 126 MB of images: a step's inputs alone exceed the 126 MB L2)
    n_batches = 3
    host = [synthetic_batch(B, seed=1000 + 17 * rank + i) for i in range(n_batches)]
    resident = [tuple(t.to(dev) for t in b) for b in host]
    pinned = [tuple(t.pin_memory() for t in b) for b in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing -------------------------------------------------------------------
    # the clock sampler (an nvidia-smi subprocess) is started BEFORE the warm-up: its start-up stalls driver
    # calls for ~100 ms, which must not land inside the timed region; it keeps sampling through it
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        train_step(model, opt, *resident[i % n_batches])
    barrier()
    if rank == 0:
        sampler.rows.clear()
    ops.timers_enable(True)
    launches0 = _lib.lib().rn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        loss = train_step(model, opt, *resident[i % n_batches])
    e1.record()
    barrier()
    launches = _lib.lib().rn_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    rel_ms = ops.timers_collect()           # {'relation_fwd': [...], 'relation_bwd': [...]}
    ops.timers_enable(False)
    final_loss = float(loss.detach())

    # ---- end to end: pinned host inputs copied every step, loss read back every step ----------------
    copy_stream = torch.cuda.Stream()
    bufs = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def stage(i):
        slot = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            for dst, src in zip(bufs[slot], pinned[i % n_batches]):
                dst.copy_(src, non_blocking=True)
            ready[slot].record(copy_stream)

    # the loss of every step is copied to pinned host memory on the compute stream and read one step later, so the
    # host never stalls the launch queue (the reference reads loss.data[0] every step, train.py:51)
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def e2e_loop(n):
        losses = []
        stage(0)
        for i in range(n):
            if i + 1 < n:
                stage(i + 1)                 # overlaps the next batch's H2D with this step's compute
            slot = i % 2
            torch.cuda.current_stream().wait_event(ready[slot])
            l = train_step(model, opt, *bufs[slot])
            consumed[slot].record()
            loss_host[slot].copy_(l.detach(), non_blocking=True)      # device -> host read of the step's result, every step
            loss_ready[slot].record()
            if i > 0:
                loss_ready[1 - slot].synchronize()
                losses.append(float(loss_host[1 - slot]))
        loss_ready[(n - 1) % 2].synchronize()
        losses.append(float(loss_host[(n - 1) % 2]))
        return losses

    for ev in consumed:
        ev.record()
    e2e_loop(max(2, args.warmup // 2))
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)

    times = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = (float(x) for x in times.cpu())

    if rank == 0:
        peaks = load_peaks()
        total_q = B * world * args.steps
        value = total_q / (ms / 1e3)
        e2e_value = total_q / (e2e_ms / 1e3)
        fwd_ms = statistics.mean(rel_ms["relation_fwd"]) if rel_ms.get("relation_fwd") else None
        bwd_ms = statistics.mean(rel_ms["relation_bwd"]) if rel_ms.get("relation_bwd") else None
        g_ms = (fwd_ms or 0.0) + (bwd_ms or 0.0)
        achieved = (G_FLOP_TRAIN * B) / (g_ms / 1e3) / 1e12 if g_ms > 0 else None
        peak = peaks["bf16_tflops_sustained"]
        h2d = sum(t.numel() * t.element_size() for t in host[0])
        line = {
            "metric": "questions/sec", "value": value, "unit": "questions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": {"fp32": "f32", "parity": "f16x2-split/f32-accum", "fast": "f16/f32-accum"}[precision],
            "data": "synthetic",
            "config": {"workload": "original-fp training step (fwd+bwd+allreduce+clip+Adam), 128x128x3 images, 8x8x24 grid, "
                                   "4096 pairs/sample, q_dim 128, batch 640 per GPU",
                       "global_batch": B * world, "per_gpu_batch": B, "precision_mode": precision,
                       "l2_policy": f"{n_batches} rotating input batches of {h2d / 1e6:.0f} MB each (> 126 MB L2); "
                                    "g-MLP activations stream through HBM-sized buffers",
                       "parallelism": f"dp{world}"},
            "e2e": {"value": e2e_value, "unit": "questions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         # DRAM bytes of one forward+backward launch pair at B=640 (parity mode), from the committed ncu
                         # capture profiles/r01c_chain_wgrad_ncu_full.txt: chain fwd 3.13 GB + dgrad 4.30 GB + wgrad
                         # 1.45 + 2.71 + 1.45 GB (+ dZ1 reduce 1.34 GB).  Far above the ~6 MB algorithmic bytes by design:
                         # training streams fp16 tile images (H2, H3, dZ1..dZ3) through HBM for the weight-gradient GEMMs
                         # (DESIGN.md section 4).
                         "traffic": 14.4e9 if (precision == "parity" and B == 640) else None,
                         "kernel": "g-MLP (rn_relation_fwd + rn_relation_bwd launches)",
                         "relation_fwd_ms": fwd_ms, "relation_bwd_ms": bwd_ms,
                         "algorithmic_flop_per_launch_pair": G_FLOP_TRAIN * B, "peak_source": peaks["source"] + " (sustained bf16)"},
            "final_loss": final_loss,
            "op_ms": {k: statistics.mean(v) for k, v in rel_ms.items()},
        }
        if args.cpu_baseline:
            qps, cores, best, n_timed = cpu_port_qps(64, 20, 2, max_seconds=25.0)          # ~10 s of CPU work on the GPU box
            line["cpu_baseline"] = {"value": qps, "unit": "questions/s", "cores": cores, "cpu": cpu_model_name(), "kind": "port",
                                    "sample": f"{n_timed} timed full training steps at batch 64 of the batch-640 workload (oracle port, "
                                              "PyTorch CPU, materialised pairs), best step"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=640)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "parity", "fast"])
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
