#!/usr/bin/env python
"""Writes profiles/ncu_metrics.json from the committed `ncu --set full` text summaries (profiles/*_ncu_full_*.txt):
per capture the metrics bench.py quotes beside its live timings (DRAM bytes of the launch, tensor-pipe activity,
duration under ncu), so that nothing in bench.py is a pasted literal.  Re-run after adding a capture:

    python tools/ncu_extract.py            # rewrites profiles/ncu_metrics.json

Each entry is keyed by the capture file; "role" says which bench.py roofline leg uses it."""
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "%": 1.0}
WANT = {"dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes",
        "gpu__time_duration.sum": "duration_us_under_ncu",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_active_pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum": "l2_to_sm_read_bytes",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct"}
# which roofline leg of bench.py a capture documents: (role, engine mode)
ROLES = {"conv2_kw_3xtf32": ("conv2_fwd", 2), "conv2_kw_tf32": ("conv2_fwd", 1), "gemm_264x512x512": ("lin_gemm", 2),
         "attn_small": ("attn_small_fwd", 2), "conv4_kw_3xtf32": ("conv4_fwd", 2), "conv2_wgrad_3xtf32": ("conv2_wgrad", 2)}


def parse(path):
    out = {}
    for line in open(path, errors="replace"):
        m = re.match(r"^\s*(\S+)\s+(\S+)\s+([0-9.,eE+-]+)\s*$", line)
        if m and m.group(1) in WANT and m.group(2) in UNIT:
            out[WANT[m.group(1)]] = float(m.group(3).replace(",", "")) * UNIT[m.group(2)]
    return out


def main():
    res = {}
    for f in sorted(os.listdir(PROF)):
        if "ncu_full" not in f or not f.endswith(".txt"):
            continue
        d = parse(os.path.join(PROF, f))
        if not d:
            continue
        key = f.split("ncu_full_")[1][:-4]
        role, mode = ROLES.get(key, (key, None))
        d.update(role=role, gemm_mode=mode, source="profiles/" + f)
        if "dram_read_bytes" in d:
            d["dram_bytes"] = d["dram_read_bytes"] + d.get("dram_write_bytes", 0.0)
        res[f] = d
    # round-2 captures: one `ncu --set full --csv --page raw` run over a whole pass / over every operator once, reduced by
    # tools/ncu_summary.py to one record per (kernel, grid)
    KERNEL_ROLES = {("conv3x3_kw_kernel<64, 1>", 1144): ("conv2_fwd", 2), ("conv3x3_kw_kernel<64, 0>", 1144): ("conv2_fwd", 1),
                    ("attn_small_fwd_mma_kernel<64, 512>", 64): ("attn_small_fwd", 2),
                    ("attn_small_bwd_mma_kernel<64, 512>", 64): ("attn_small_bwd", 2),
                    ("ln_fwd_kernel", 66): ("ln_fwd", 2), ("ln_bwd_kernel", 66): ("ln_bwd", 2)}
    for f in sorted(os.listdir(PROF)):
        if not f.endswith("_kernels.json"):
            continue
        for rec in json.load(open(os.path.join(PROF, f))):
            name = rec["kernel"].replace("void ", "").replace("<unnamed>::", "")
            key = (name, int(rec.get("grid") or 0))
            role, mode = KERNEL_ROLES.get(key, (None, None))
            if role == "conv2_fwd" and "forward" not in f:         # same kernel and grid serve conv.2's input gradient
                role = "conv2_dgrad"
            d = {"role": role or name, "gemm_mode": mode, "source": "profiles/" + f, "kernel": name, "grid": key[1],
                 "duration_us_under_ncu": rec.get("dur_us"), "dram_read_bytes": rec.get("dram_rd"),
                 "dram_write_bytes": rec.get("dram_wr"), "dram_bytes": (rec.get("dram_rd") or 0) + (rec.get("dram_wr") or 0),
                 "tensor_pipe_active_pct": rec.get("tensor_pct"), "sm_throughput_pct": rec.get("sm_pct"),
                 "l2_to_sm_read_bytes": rec.get("l2_to_sm")}
            res["%s::%s@%d" % (f, name, key[1])] = d
    with open(os.path.join(PROF, "ncu_metrics.json"), "w") as fh:
        json.dump(res, fh, indent=1, sort_keys=True)
    print("wrote %d captures" % len(res))


if __name__ == "__main__":
    main()
