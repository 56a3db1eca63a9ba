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


if __name__ == "__main__" and sys.argv[1] == "laststep":
    launches(sys.argv[2], int(sys.argv[3]))
    sys.exit(0)

if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
