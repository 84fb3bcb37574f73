"""Key `ncu --set full` metrics of one or more reports as a markdown table.
Usage: python tools/ncu_raw_summary.py name=report.ncu-rep [name=report ...]"""
import csv, io, subprocess, sys
METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("sm__cycles_elapsed.avg", "SM cycles"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "tensor pipe (HMMA) active cycles"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "TMA (mem tensor) unit active %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "LSU shared-memory wavefronts % of peak"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts (LSU)"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.sum.per_cycle_elapsed", "IPC (per SM)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
]
cols = []
for arg in sys.argv[1:]:
    name, rep = arg.rsplit("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    cols.append((name, {h: (v, u) for h, u, v in zip(hdr, units, vals)}))
print("| metric | " + " | ".join(n for n, _ in cols) + " |")
print("|---|" + "---|" * len(cols))
for key, label in METRICS:
    cells = []
    for _, d in cols:
        hit = [(k, v) for k, v in d.items() if k.endswith(key)]
        if hit:
            v, u = hit[0][1]
            try:
                v = "%.4g" % float(v)
            except ValueError:
                pass
            cells.append("%s %s" % (v, u) if u else v)
        else:
            cells.append("")
    print("| `%s` (%s) | " % (key, label) + " | ".join(cells) + " |")
