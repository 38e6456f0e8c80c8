"""Summarise an `ncu --set full` report per kernel (first captured launch of each name): duration, DRAM bytes, pipe
utilisation, issue rate, occupancy.  usage: python tools/ncu_summary.py report.ncu-rep [attention_json_out]"""
import csv
import io
import json
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
ix = {h: i for i, h in enumerate(hdr)}
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "sm__warps_active.avg.pct_of_peak_sustained_active"]
seen = set()
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ix["Kernel Name"]]).replace("mmpfn::<unnamed>::", "").replace("unnamed>::", "").replace("void ", "")
    if name in seen:
        continue
    seen.add(name)
    print(name)
    vals = {}
    for m in METRICS:
        if m in ix:
            v = r[ix[m]].replace(",", "")
            vals[m] = v
            print(f"    {m:75s} {float(v):16.3f} {units[ix[m]]}")
    t_us = float(vals["gpu__time_duration.sum"]) * (1e-3 if units[ix["gpu__time_duration.sum"]] in ("nsecond", "ns") else 1.0)
    def to_bytes(m):
        v, u = float(vals[m]), units[ix[m]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    print(f"    -> DRAM {(rd + wr) / 1e6:.1f} MB in {t_us:.1f} us = {(rd + wr) / t_us / 1e6:.2f} TB/s\n")
    if len(sys.argv) > 2 and "tc_item_attn" in name:
        json.dump({"kernel": name, "shape": {"B": 4, "T": 27, "n_q": 2000, "n_kv": 2000}, "dram_bytes_read": rd,
                   "dram_bytes_write": wr, "gpu_time_us_under_ncu": t_us,
                   "xu_pipe_pct": float(vals["sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]),
                   "tensor_pipe_pct": float(vals["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]),
                   "issue_active_pct": float(vals["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
                   "instructions": float(vals["smsp__inst_executed.sum"]),
                   "source": "profiles/<the summary written next to this json> (ncu --set full --clock-control none, one launch, tools/prof_kernels.py)"},
                  open(sys.argv[2], "w"), indent=1)
