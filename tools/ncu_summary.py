"""Summarise ncu outputs into profiles/: (a) a launch list CSV -> per-kernel totals and shares, (b) an .ncu-rep -> the
handful of metrics DESIGN.md quotes.   python tools/ncu_summary.py launches <csv> | full <ncu-rep>"""
import collections
import csv
import re
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[[i for i, r in enumerate(rows) if r and r[0] == "ID"][0]]
    data = rows[rows.index(h) + 1:]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in data:
        n = re.sub(r"\(.*", "", r[ki]).split("::")[-1]
        v = float(r[vi].replace(",", "")) / (1e6 if r[ui] == "ns" else 1e3 if r[ui] == "us" else 1.0)
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f"# per-kernel device time from {path} (ncu gpu__time_duration.sum, cold-cache serialised launches: compare SHARES)")
    print(f"# total {tot:.3f} ms over {sum(a[0] for a in agg.values())} launches")
    for n, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{n:36s} launches {c:4d}  time {t:10.3f} ms  share {100 * t / tot:6.2f}%")


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor_subpipe_hmma",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_read.sum", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:                      # one block per captured launch
        if len(vals) != len(hdr):
            continue
        print(f"# ncu --set full --clock-control none, kernel {vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else ''}")
        for h, u, v in zip(hdr, units, vals):
            if any(h == k or h.startswith(k + " ") or (k in h and h.endswith(k.split('.')[-1])) for k in KEYS) and any(h.startswith(k) for k in KEYS):
                print(f"{h} [{u}] = {v}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
