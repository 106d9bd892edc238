#!/usr/bin/env python
"""Summarise an `ncu --set full` report into the small JSON that profiles/ keeps and bench.py reads.

    ncu -i gpurun_out/rNN_full.ncu-rep --page raw --csv > /tmp/raw.csv
    python tools/ncu_summary.py /tmp/raw.csv profiles/rNN_ncu_full_kernels.json
"""
import csv
import json
import re
import sys

FIELDS = {
    "time": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "dram_read_pct": "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
    "dram_write_pct": "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
    "l2_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dmma_pct_active": "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "dyn_smem": "launch__shared_mem_per_block_dynamic",
    "smem_bank_conflicts": "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
}


def main(src, dst):
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in data:
        name = re.sub(r"^void\s+", "", r[col["Kernel Name"]])
        name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name).split("(")[0]
        e = {"kernel": name}
        for key, metric in FIELDS.items():
            if metric in col:
                e[key] = f"{r[col[metric]]} {units[col[metric]]}".strip()
        out.append(e)
    json.dump(out, open(dst, "w"), indent=1)
    for e in out:
        print(e["kernel"][:40].ljust(40), e.get("time"), "| DRAM", e.get("dram_read"), "+", e.get("dram_write"), "|", e.get("dram_read_pct"), "+", e.get("dram_write_pct"), "| dmma", e.get("dmma_pct_active"))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
