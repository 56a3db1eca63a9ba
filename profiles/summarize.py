#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python profiles/summarize.py launches gpurun_out/launches.csv > profiles/r01_launches.txt
  python profiles/summarize.py full gpurun_out/prof.ncu-rep      > profiles/r01_<kernel>_ncu.txt
"""
import collections
import csv
import subprocess
import sys

KEEP = ("gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "sm__warps_active.avg.per_cycle_active", "launch__registers_per_thread",
        "sm__cycles_elapsed.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed")


def launches(path, steps=1):
    rows = list(csv.reader(open(path, errors="ignore")))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    body = [r for r in rows[hi + 1:] if len(r) > mv]
    if steps > 1:                       # the capture holds `steps` identical steps: keep the last one
        body = body[len(body) - len(body) // steps:]
    for r in body:
        if len(r) <= mv:
            continue
        a = agg.setdefault(r[kn][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[mv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.3f} ms total device time (ncu, serialised)")
    for n, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1] / 1e6:9.3f} ms {a[0]:5d}x {100 * a[1] / tot:5.1f}%  {n}")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("KERNEL", r[hdr.index("Kernel Name")])
        for i, h in enumerate(hdr):
            if h in KEEP:
                print(f"   {h:85s} {units[i]:12s} {r[i]}")


def steprows(path):
    """One line per kernel of a `--set full` capture of one training step (profiles/r02*_step_b640_ncu_full.txt)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}

    def val(r, k):
        return float(r[col[k]].replace(",", "") or 0)

    print("# columns: kernel | time ms | tensor pipe % | DRAM read GB | DRAM write GB | DRAM % of peak | achieved DRAM GB/s (vs 6553 measured copy)"
          " | L2 % | issue active % | regs | grid x block")
    tot = 0.0
    for r in rows[2:]:
        t = val(r, "gpu__time_duration.sum") * tscale[units[col["gpu__time_duration.sum"]]]
        rd = val(r, "dram__bytes_read.sum") * scale[units[col["dram__bytes_read.sum"]]] / 1e9
        wr = val(r, "dram__bytes_write.sum") * scale[units[col["dram__bytes_write.sum"]]] / 1e9
        name = r[col["Kernel Name"]].split("(")[0].replace("rn::", "")
        tot += t
        print(f"{name[:44]:44s} {t:6.3f} ms  tensor {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f}%"
              f"  rd {rd:6.3f} GB  wr {wr:6.3f} GB  dram {val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}%"
              f"  {(rd + wr) / (t / 1e3) if t else 0:7.0f} GB/s  L2 {val(r, 'lts__throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}%"
              f"  issue {val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active') if 'smsp__issue_active.avg.pct_of_peak_sustained_active' in col else float('nan'):5.1f}%"
              f"  regs {int(val(r, 'launch__registers_per_thread'))}  {int(val(r, 'launch__grid_size'))}x{int(val(r, 'launch__block_size'))}")
    print(f"# {len(rows) - 2} kernels, {tot:.3f} ms serialised")


if __name__ == "__main__" and sys.argv[1] == "steprows":
    steprows(sys.argv[2])
    sys.exit(0)

if __name__ == "__main__" and sys.argv[1] == "laststep":
    launches(sys.argv[2], int(sys.argv[3]))
    sys.exit(0)

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
