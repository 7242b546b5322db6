#!/usr/bin/env python
"""`ncu --set full --csv --page raw` of tools/probes/one_pass.py -> one row per distinct kernel (first launch of each
(name, grid)): duration, DRAM bytes, DRAM / tensor-pipe / SM utilisation, occupancy, registers, shared memory.
usage: python tools/ncu_summary.py raw.csv [out.json] > profiles/rNN_ncu_pass_kernels.txt"""
import csv
import json
import sys

COLS = {"gpu__time_duration.sum": "dur_us", "dram__bytes_read.sum": "dram_rd", "dram__bytes_write.sum": "dram_wr",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "occ_pct", "launch__registers_per_thread": "regs",
        "launch__shared_mem_per_block_dynamic": "smem_dyn", "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm",
        "launch__grid_size": "grid", "launch__block_size": "block"}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6,
         "ns": 1e-3, "us": 1.0, "ms": 1e3}


def num(v, unit):
    try:
        return float(v.replace(",", "")) * SCALE.get(unit, 1.0)
    except ValueError:
        return None


def main(path, out_json=None):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, units, data = rows[hi], rows[hi + 1], rows[hi + 2:]
    ki = hdr.index("Kernel Name")
    idx = {c: hdr.index(c) for c in COLS if c in hdr}
    seen, out = set(), []
    for r in data:
        if len(r) <= ki:
            continue
        name = r[ki].replace("(anonymous namespace)::", "").split("(")[0]
        rec = {COLS[c]: num(r[i], units[i]) for c, i in idx.items()}
        key = (name, rec.get("grid"))
        if key in seen:
            continue
        seen.add(key)
        rec["kernel"] = name
        out.append(rec)
    print("# %s: %d launches, %d distinct (kernel, grid) pairs; per-launch values, cold cache, serialised" % (path, len(data), len(out)))
    print("%-64s %8s %9s %10s %10s %7s %7s %6s %6s %5s %8s" % ("kernel", "grid", "dur us", "DRAM rd MB", "DRAM wr MB", "DRAM %", "tensor%", "SM %", "occ %", "regs", "smem KB"))
    f = lambda v, s="%.1f": "-" if v is None else s % v
    for r in sorted(out, key=lambda r: -(r.get("dur_us") or 0)):
        print("%-64s %8s %9s %10s %10s %7s %7s %6s %6s %5s %8s" % (
            r["kernel"][-64:], f(r.get("grid"), "%d"), f(r.get("dur_us"), "%.2f"), f((r.get("dram_rd") or 0) / 1e6, "%.3f"),
            f((r.get("dram_wr") or 0) / 1e6, "%.3f"), f(r.get("dram_pct")), f(r.get("tensor_pct")), f(r.get("sm_pct")),
            f(r.get("occ_pct")), f(r.get("regs"), "%d"), f((r.get("smem_dyn") or 0) / 1e3)))
    if out_json:
        json.dump(out, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(*sys.argv[1:3])
