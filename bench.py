#!/usr/bin/env python
"""Headline benchmark: questions/s of one Relation-Network TRAINING step (forward + backward + gradient all-reduce +
clip/Adam) on BASELINE.json config 2 / 3: original-fp, 128x128 images, 8x8x24 grid, GLOBAL batch 640, synthetic data,
seeded random-init weights.

    python bench.py --gpus 1 --steps 30 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the UNMODIFIED reference model.py (baseline/_ref) on the host cores
    python bench.py --config ir-fp            # BASELINE config 4
    python bench.py --grid 12 | --grid 16     # BASELINE config 5 (image side 16 d, constant pair-row count batches)

Prints ONE JSON line (rank 0).  N > 1: the headline `value` is STRONG scaling -- the DataParallel-equivalent reading of the
reference (train.py:256-258 splits ONE batch of 640 over the GPUs: 640/N questions per GPU) -- and the same line carries
the weak-scaling measurement (640 per GPU) under "weak".  `value` = device-resident throughput (CUDA-graphed step);
`e2e` = the same step through the public API with pinned-host inputs copied every step and the loss read back;
`roofline` = the g-MLP (relation forward + backward launches) against the measured bf16 tensor peak; `cpu_baseline` =
the reference timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import contextlib
import io
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
GLOBAL_BATCH = 640
# algorithmic g-MLP FLOPs per pair (SURVEY.md 8d / BASELINE.md 2): 242,688 MAC x 2, x 3 for forward + backward
G_FLOP_PAIR_FWD = 2.0 * 242_688
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def measured_traffic(key: str):
    """DRAM bytes of one g-MLP forward + backward launch pair from the committed ncu capture (profiles/, written by
    profiles/summarize.py from `ncu --set full`), or None when no capture exists for this configuration."""
    path = os.path.join(ROOT, "profiles", "r02_gmlp_traffic.json")
    try:
        return json.load(open(path)).get(key, {}).get("dram_bytes_per_launch_pair")
    except (OSError, ValueError):
        return None


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


def synthetic_batch(B: int, side: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, 3, side, side, generator=g)                       # ToTensor range, no normalisation
    qst = torch.randint(1, QDICT + 1, (B, T_Q), generator=g, dtype=torch.int64)
    lab = torch.randint(0, ADICT, (B,), generator=g, dtype=torch.int64)
    return img, qst, lab


def workload_name(args) -> str:
    d = args.grid
    return (f"{args.config} training step (fwd+bwd+allreduce+clip+Adam), {16 * d}x{16 * d}x3 images, {d}x{d}x24 grid, "
            f"{d ** 4} pairs/sample, q_dim 128, global batch {args.batch}")


# --------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's own CPU implementation of the path on the host cores.
# build() stages the UNMODIFIED /root/reference/model.py under baseline/_ref/ (git-ignored; it travels to the GPU box
# with the repo snapshot); if that file is missing the oracle port (oracle/rn_oracle.py, pinned to the reference by
# tests/golden) is timed instead and the line says kind = "port".
# --------------------------------------------------------------------------------------------------
def cpu_model_name() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def _reference_stepper(config: str, B: int, side: int):
    """(kind, step_fn): one full training step (zero_grad, forward, nll_loss, backward, clip_grad_norm 50, Adam wd 1e-4)."""
    hyp = json.load(open(os.path.join(ROOT, "config.json")))["hyperparams"][config]
    img, qst, lab = synthetic_batch(B, side, seed=42)
    if os.path.exists(os.path.join(REF_DIR, "model.py")):
        sys.path.insert(0, REF_DIR)
        import model as refmodel      # the reference's model.py, byte for byte

        class A:
            qdict_size, adict_size = QDICT, ADICT

        torch.manual_seed(42)
        with contextlib.redirect_stdout(io.StringIO()):
            m = refmodel.RN(A, hyp)
        m.train()
        opt = torch.optim.Adam(m.parameters(), lr=5e-6, weight_decay=1e-4)        # train.py:330

        def step():
            opt.zero_grad()
            loss = F.nll_loss(m(img, qst), lab)                                    # train.py:40-41
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 50)                     # train.py:45
            opt.step()
            return float(loss.detach())

        return "reference", step
    from oracle import rn_oracle as O

    p = O.seeded_params(O.HYPERPARAMS[config], QDICT, ADICT, seed=42)
    names = [k for k in p if "running" not in k]
    leaves = {k: (p[k].clone().requires_grad_(True) if k in names else p[k].clone()) for k in p}
    plist = [leaves[k] for k in names]
    m1 = [torch.zeros_like(w) for w in plist]
    m2 = [torch.zeros_like(w) for w in plist]
    it = [0]

    def step():
        for w in plist:
            w.grad = None
        loss = F.nll_loss(O.rn_forward(leaves, O.HYPERPARAMS[config], img, qst, training=True), lab)
        loss.backward()
        it[0] += 1
        with torch.no_grad():
            O.clip_and_adam(plist, [w.grad for w in plist], m1, m2, it[0], lr=5e-6)
        return float(loss.detach())

    return "port", step


def cpu_reference_qps(config: str, side: int, batch: int, steps: int, warmup: int, max_seconds: float):
    """Times full training steps of the reference on the host; tries the full batch first and falls back to a batch-64
    sample if the box cannot hold it (the literal pair tensors of batch 640 need ~19 GB)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    last_err = None
    for B in (batch, 64):
        try:
            kind, step = _reference_stepper(config, B, side)
            times = []
            t_start = time.perf_counter()
            for it in range(warmup + steps):
                if len(times) >= 1 and time.perf_counter() - t_start > max_seconds:
                    break                                  # bounded sample
                t0 = time.perf_counter()
                step()
                if it >= warmup:
                    times.append(time.perf_counter() - t0)
            med = statistics.median(times)
            return {"qps": B / med, "cores": cores, "kind": kind, "B": B, "median_s": med, "timed_steps": len(times)}
        except (RuntimeError, MemoryError) as e:          # out of host memory at the full batch
            last_err = e
    raise last_err


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    side = 16 * args.grid
    steps = max(1, min(args.steps, 3))            # ~5-15 s per batch-640 step on 16 host cores
    warm = 1
    r = cpu_reference_qps(args.config, side, args.batch, steps, warm, max_seconds=170.0)
    sample = (f"{r['timed_steps']} timed full training steps (fwd+bwd+clip_grad_norm+Adam) of "
              f"{'the unmodified reference model.py' if r['kind'] == 'reference' else 'the oracle port'} at batch {r['B']}, "
              f"{warm} warm-up, median step, torch {torch.__version__} CPU, {r['cores']} threads")
    line = {"impl": "reference", "metric": "questions/sec", "value": r["qps"], "unit": "questions/s", "n_gpus": args.gpus,
            "steps": r["timed_steps"], "warmup": warm, "ms_per_step": r["median_s"] * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": r["B"]},
            "cpu_baseline": {"value": r["qps"], "unit": "questions/s", "cores": r["cores"], "cpu": cpu_model_name(),
                             "kind": r["kind"], "sample": sample},
            "e2e": {"value": r["qps"], "unit": "questions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_ours(args):
    import relationnetworks_clevr_b200 as R
    from relationnetworks_clevr_b200 import _lib, ops
    from relationnetworks_clevr_b200.trainer import FlatClipAdam, GraphedTrainStep, train_step

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
    if args.batch % world:
        raise SystemExit(f"global batch {args.batch} is not divisible by {world} ranks")
    side = 16 * args.grid
    n_obj = args.grid ** 2

    class A:
        qdict_size, adict_size = QDICT, ADICT

    hyp = json.load(open(os.path.join(ROOT, "config.json")))["hyperparams"][args.config]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(B: int, want_e2e: bool, want_clocks: bool):
        """One full measurement at `B` questions per GPU: device-resident (graphed) and, optionally, end to end."""
        ops.clear_grad_sink()
        torch.manual_seed(42)                       # train.py:382 -- identical random init on all ranks
        with contextlib.redirect_stdout(io.StringIO()):
            model = R.RN(A, hyp)
        model.to(dev).train()
        model.rl.precision = args.precision
        opt = FlatClipAdam(model.parameters(), lr=5e-6, weight_decay=1e-4, clip_norm=50.0)
        precision = model.rl._resolve_precision(n_obj, 26)
        # several distinct device-resident batches (126 MB of images at 640 x 128 x 128: a step's inputs alone match the
        # 126 MB L2, and the g-MLP streams GBs of activations through HBM in between)
        n_batches = 3
        host = [synthetic_batch(B, side, seed=1000 + 17 * rank + i) for i in range(n_batches)]
        resident = [tuple(t.to(dev) for t in b) for b in host]
        sampler = ClockSampler(local_rank)
        if rank == 0 and want_clocks:
            sampler.start()              # before the warm-up: its start-up stalls driver calls for ~100 ms
        # launches of this library per step, counted on an eager step (graph replays do not pass through the C ABI)
        train_step(model, opt, *resident[0])
        l0 = _lib.lib().rn_launch_count()
        train_step(model, opt, *resident[1])
        launches_per_step = int(_lib.lib().rn_launch_count() - l0)
        graphed = GraphedTrainStep(model, opt, *resident[0]) if not args.no_graph else None
        use_graph = graphed is not None and graphed.captured
        run = (lambda b: graphed.step(*b)) if use_graph else (lambda b: train_step(model, opt, *b))
        for i in range(args.warmup):
            run(resident[i % n_batches])
        barrier()
        if rank == 0 and want_clocks:
            sampler.rows.clear()
        # ---- device-resident timing: K steps between two events, plus one event per step for the median ----
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        marks[0].record()
        for i in range(args.steps):
            loss = run(resident[i % n_batches])
            marks[i + 1].record()
        barrier()
        clocks = sampler.stop() if (rank == 0 and want_clocks) else None
        ms = marks[0].elapsed_time(marks[-1])
        per_step = [marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps)]
        final_loss = float(loss.detach())
        # ---- per-op timers (eager: CUDA events around each C-ABI call), a few steps ----
        ops.timers_enable(True)
        for i in range(min(args.steps, 8)):
            train_step(model, opt, *resident[i % n_batches])
        op_ms = {k: statistics.median(v) for k, v in ops.timers_collect().items()}
        ops.timers_enable(False)
        out = {"B": B, "ms": ms, "per_step": per_step, "final_loss": final_loss, "op_ms": op_ms, "clocks": clocks,
               "precision": precision, "launches_per_step": launches_per_step, "graph": use_graph,
               "h2d": sum(t.numel() * t.element_size() for t in host[0])}
        if want_e2e:
            # ---- end to end: pinned host inputs copied every step (copy stream, double-buffered staging), loss read back ----
            pinned = [tuple(t.pin_memory() for t in b) for b in host]
            copy_stream = torch.cuda.Stream()
            bufs = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]
            loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
            loss_ready = [torch.cuda.Event() for _ in range(2)]

            def stage(i):
                slot = i % 2
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[slot])
                    for dst, src in zip(bufs[slot], pinned[i % n_batches]):
                        dst.copy_(src, non_blocking=True)
                    ready[slot].record(copy_stream)

            def e2e_loop(n):
                # the loss of every step is copied to pinned host memory on the compute stream and read one step later, so
                # the host never stalls the launch queue (the reference reads loss.data[0] every step, train.py:51)
                losses = []
                stage(0)
                for i in range(n):
                    if i + 1 < n:
                        stage(i + 1)                 # overlaps the next batch's H2D with this step's compute
                    slot = i % 2
                    torch.cuda.current_stream().wait_event(ready[slot])
                    l = run(bufs[slot])              # graphed: device-to-device copy into the static buffers, then replay
                    consumed[slot].record()
                    loss_host[slot].copy_(l.detach(), non_blocking=True)
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
            out["e2e_ms"] = t0.elapsed_time(t1)
            # ---- the same, staging RAW uint8 pixels (the first conv layer divides by 255 while staging: f3 of SURVEY 8f) ----
            host_u8 = [((b[0] * 255).round().to(torch.uint8), b[1], b[2]) for b in host]
            pinned = [tuple(t.pin_memory() for t in b) for b in host_u8]
            bufs = [tuple(torch.empty_like(t, device=dev) for t in host_u8[0]) for _ in range(2)]
            g8 = GraphedTrainStep(model, opt, host_u8[0][0].to(dev), resident[0][1], resident[0][2]) if use_graph else None
            run = (lambda b: g8.step(*b)) if (g8 is not None and g8.captured) else (lambda b: train_step(model, opt, *b))
            for ev in consumed:
                ev.record()
            e2e_loop(max(2, args.warmup // 2))
            barrier()
            t0.record()
            e2e_loop(args.steps)
            t1.record()
            barrier()
            out["e2e_u8_ms"] = t0.elapsed_time(t1)
            out["h2d_u8"] = sum(t.numel() * t.element_size() for t in host_u8[0])
            del g8
        del graphed, model, opt
        ops.clear_grad_sink()
        torch.cuda.empty_cache()
        return out

    def reduce_max(vals):
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [float(x) for x in t.cpu()]

    strong = measure(args.batch // world, want_e2e=True, want_clocks=True)
    s_ms, s_e2e, s_e2e_u8 = reduce_max([strong["ms"], strong["e2e_ms"], strong["e2e_u8_ms"]])
    weak = None
    if world > 1 and not args.no_weak:
        weak = measure(args.batch, want_e2e=False, want_clocks=False)
        (w_ms,) = reduce_max([weak["ms"]])

    if rank == 0:
        peaks = load_peaks()
        B = strong["B"]
        total_q = B * world * args.steps
        value = total_q / (s_ms / 1e3)
        e2e_value = total_q / (s_e2e / 1e3)
        op_ms = strong["op_ms"]
        fwd_ms, bwd_ms = op_ms.get("relation_fwd"), op_ms.get("relation_bwd")
        flop_pair = 3.0 * G_FLOP_PAIR_FWD * (n_obj ** 2) * B
        g_ms = (fwd_ms or 0.0) + (bwd_ms or 0.0)
        achieved = flop_pair / (g_ms / 1e3) / 1e12 if g_ms > 0 else None
        peak = peaks["bf16_tflops_sustained"]
        precision = strong["precision"]
        dtype = {"fp32": "f32 (SIMT)",
                 "parity": "f16 operands / f32 accumulate: forward 3 MMA passes (A_hi W_hi + A_lo W_hi + A_hi W_lo), "
                           "data gradient 1 pass (W_hi), weight gradient 1 pass; conv backward TF32 operands x 3 passes (hi/lo split, "
                           "f32-level), f32 accumulate; layer 0, conv forward, LSTM, f-MLP, Adam in f32",
                 "fast": "f16 operands / f32 accumulate, 1 MMA pass everywhere"}[precision]
        line = {
            "metric": "questions/sec", "value": value, "unit": "questions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": s_ms / args.steps, "ms_per_step_median": statistics.median(strong["per_step"]),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": {"workload": workload_name(args), "global_batch": B * world, "per_gpu_batch": B,
                       "precision_mode": precision, "cuda_graph": strong["graph"],
                       "l2_policy": f"3 rotating input batches of {strong['h2d'] / 1e6:.0f} MB each per GPU; the g-MLP streams "
                                    f"{'GBs' if B >= 160 else 'hundreds of MB'} of activation images through HBM between them",
                       "parallelism": f"dp{world}"},
            "e2e": {"value": e2e_value, "unit": "questions/s", "h2d_bytes_per_step": strong["h2d"], "d2h_bytes_per_step": 4,
                    "ms_per_step": s_e2e / args.steps},
            "e2e_u8": {"value": total_q / (s_e2e_u8 / 1e3), "unit": "questions/s", "h2d_bytes_per_step": strong["h2d_u8"],
                       "d2h_bytes_per_step": 4, "ms_per_step": s_e2e_u8 / args.steps,
                       "note": "same step fed raw uint8 pixels (converted u/255 inside the first conv layer, bit-identical to "
                               "ToTensor on the host); `e2e` above stages the fp32 tensors the reference's loader produces"},
            "gpu_launches": strong["launches_per_step"] * args.steps,
            "gpu_launches_per_step": strong["launches_per_step"],
            "clocks": strong["clocks"],
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": measured_traffic(f"{args.config}_d{args.grid}_b{B}_{precision}"),
                         "kernel": "g-MLP (rn_relation_fwd + rn_relation_bwd launches)",
                         "relation_fwd_ms": fwd_ms, "relation_bwd_ms": bwd_ms,
                         "algorithmic_flop_per_launch_pair": flop_pair, "peak_source": peaks["source"] + " (sustained bf16)"},
            "final_loss": strong["final_loss"],
            "op_ms": op_ms,
        }
        if weak is not None:
            line["weak"] = {"value": weak["B"] * world * args.steps / (w_ms / 1e3), "unit": "questions/s",
                            "per_gpu_batch": weak["B"], "global_batch": weak["B"] * world, "ms_per_step": w_ms / args.steps,
                            "ms_per_step_median": statistics.median(weak["per_step"]), "scaling": "weak"}
        if args.cpu_baseline:
            r = cpu_reference_qps(args.config, side, args.batch, 1, 1, max_seconds=40.0)
            line["cpu_baseline"] = {"value": r["qps"], "unit": "questions/s", "cores": r["cores"], "cpu": cpu_model_name(),
                                    "kind": r["kind"],
                                    "sample": f"{r['timed_steps']} timed full training step(s) at batch {r['B']} after 1 warm-up "
                                              f"({'unmodified reference model.py' if r['kind'] == 'reference' else 'oracle port'}, "
                                              "PyTorch CPU, materialised pairs)"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="original-fp", choices=["original-fp", "ir-fp"])
    ap.add_argument("--grid", type=int, default=8, choices=[8, 12, 16], help="d of the d x d object grid (image side 16 d)")
    ap.add_argument("--batch", type=int, default=None, help="GLOBAL batch (default 640; 128 / 40 for --grid 12 / 16)")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "parity", "fast"])
    ap.add_argument("--no-graph", action="store_true", help="eager steps instead of CUDA-graph replays")
    ap.add_argument("--no-weak", action="store_true", help="N > 1: skip the secondary weak-scaling measurement")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.batch is None:
        args.batch = {8: GLOBAL_BATCH, 12: 128, 16: 40}[args.grid]
    if args.gpus > 1 or args.impl == "reference":
        args.cpu_baseline = args.cpu_baseline and False       # rank 0 at N = 1 only
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
