"""Summarise the raw-page CSV of an `ncu --set full` capture (ncu -i X.ncu-rep --page raw --csv > X.raw.csv, done on the
GPU box so that only the CSV has to travel) into a small text table for profiles/.
    python tools/ncu_csv_summary.py gpurun_out/r02_ncu_matpow.raw.csv > profiles/r02_ncu_matpow.txt"""
import csv, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
MULT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
print(f"# {sys.argv[1]} — ncu --set full --clock-control none (cold-cache, serialised replays: compare shares and traffic, not absolutes)")
for r in rows[2:]:
    print("=" * 110)
    print(r[hdr.index("Kernel Name")][:108])
    rd = wr = None
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"  {k:70s} {r[i]:>16s} {units[i]}")
            if k == "dram__bytes_read.sum": rd = float(r[i].replace(",", "")) * MULT.get(units[i], 1)
            if k == "dram__bytes_write.sum": wr = float(r[i].replace(",", "")) * MULT.get(units[i], 1)
    if rd is not None and wr is not None:
        i = hdr.index("gpu__time_duration.sum")
        dur = float(r[i].replace(",", "")) * TIME.get(units[i], 1e-9)
        print(f"  {'-> DRAM traffic per launch (read + write)':70s} {(rd + wr) / 1e9:16.4f} GB   ({(rd + wr) / dur / 1e9:.0f} GB/s under ncu)")
