"""Selected metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv) as JSON: the extracts kept under profiles/."""
import csv, io, json, subprocess, sys
KEYS = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__cluster_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum"]
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
names, units = rows[hdr], rows[hdr + 1]
res = []
for r in rows[hdr + 2:]:
    if len(r) != len(names):
        continue
    d = {}
    for n, u, v in zip(names, units, r):
        if any(n == k for k in KEYS):
            d[f"{n} [{u}]"] = v
    res.append(d)
json.dump(res, sys.stdout, indent=1)
