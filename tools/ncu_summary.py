"""Summarise an `ncu --set full` report: one JSON object per captured launch with the metrics the roofline argument
needs (duration, DRAM bytes, achieved DRAM GB/s, FP64 tensor / FP64 pipe utilisation, L2 hit rate, shared-memory bank
conflicts, registers, warp-stall breakdown).
    python tools/ncu_summary.py report.ncu-rep [--extra key=value ...] > profiles/<name>_summary.json
Needs the `ncu` CLI (reads the report, no GPU)."""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
extra = dict(a.split("=", 1) for a in sys.argv[2:] if "=" in a and not a.startswith("--"))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0,
        "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0}


def val(r, name, scale=True):
    if name not in ix:
        return None
    try:
        v = float(r[ix[name]].replace(",", ""))
    except ValueError:
        return None
    return v * UNIT.get(units[ix[name]], 1.0) if scale else v


KEEP = ["sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
out = []
for r in data:
    t = val(r, "gpu__time_duration.sum")
    rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
    o = {"kernel": r[ix["Kernel Name"]], "duration_ms": None if t is None else 1e3 * t, "dram_bytes_read": rd, "dram_bytes_write": wr,
         "dram_gb_per_s": None if not t or rd is None else (rd + (wr or 0.0)) / t / 1e9}
    for k in KEEP:
        v = val(r, k, scale=False)
        if v is not None:
            o[k] = v
    stalls = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: val(r, h, scale=False) for h in hdr
              if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    tot = sum(v for v in stalls.values() if v)
    if tot:
        o["stall_samples_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda x: -(x[1] or 0)) if v and 100 * v / tot >= 0.5}
    for k, v in extra.items():
        try:
            o[k] = float(v) if "." in v else int(v)
        except ValueError:
            o[k] = v
    out.append(o)
print(json.dumps(out[0] if len(out) == 1 else out, indent=1))
