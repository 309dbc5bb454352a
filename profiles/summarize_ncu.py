"""profiles/summarize_ncu.py <report.ncu-rep> [algorithmic_bytes] -- text summary of an `ncu --set full` capture
(run here, no GPU needed: `ncu -i ... --page raw --csv`).  The summaries under profiles/ are made with it."""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]

rep = sys.argv[1]
alg = float(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
TO_BYTES = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
TO_S = {"ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0, "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}
for r in rows[2:]:
    d, u = dict(zip(hdr, r)), dict(zip(hdr, units))
    print("==", d["Kernel Name"])
    for k in KEYS:
        if k in d:
            print(f"  {k:88s} {d[k]:>18s} {u[k]}")
    for k in hdr:
        if "issue_stalled" in k and k.endswith("per_issue_active.ratio"):
            print(f"  {k:88s} {d[k]:>18s} inst")
    try:
        traffic = float(d["dram__bytes_read.sum"].replace(",", "")) * TO_BYTES[u["dram__bytes_read.sum"]] + \
            float(d["dram__bytes_write.sum"].replace(",", "")) * TO_BYTES[u["dram__bytes_write.sum"]]
        t = float(d["gpu__time_duration.sum"].replace(",", "")) * TO_S[u["gpu__time_duration.sum"]]
        msg = f"  -> dram read+write per launch: {traffic / 1e9:.3f} GB"
        if alg:
            msg += f"; algorithmic bytes {alg / 1e9:.3f} GB; achieved {alg / t / 1e9:.0f} GB/s under ncu (serialised, cold)"
        print(msg)
    except Exception as e:   # noqa: BLE001
        print("  (no traffic summary:", e, ")")
    print()
