"""Summarise an `ncu --set full` report (.ncu-rep) into a small text table for profiles/."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for r in rows[2:]:
    name = r[hdr.index("Kernel Name")]
    print("=" * 100)
    print(name[:98])
    rd = wr = None
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:68s} {r[i]:>16s} {units[i]}")
            if k == "dram__bytes_read.sum": rd = (float(r[i].replace(",", "")), units[i])
            if k == "dram__bytes_write.sum": wr = (float(r[i].replace(",", "")), units[i])
    if rd and wr:
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd[0] * mult[rd[1]] + wr[0] * mult[wr[1]]
        dur = float(r[hdr.index("gpu__time_duration.sum")].replace(",", ""))
        du = units[hdr.index("gpu__time_duration.sum")]
        dur_s = dur * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}[du]
        print(f"  {'-> DRAM traffic per launch (read+write)':68s} {tot/1e9:16.4f} GB   ({tot/dur_s/1e9:.0f} GB/s under ncu)")
