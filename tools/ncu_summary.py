"""Transposes `ncu -i X.ncu-rep --page raw --csv` (one very wide row per kernel) into the
`metric,unit,<kernel>` CSV kept under profiles/, and prints the handful of numbers DESIGN.md quotes:
duration, DRAM bytes, LSU wavefronts (shared ld / st, global side) per 32 tets, bank conflicts,
executed instructions per tet, issue / LSU pipe utilisation.

usage: python tools/ncu_summary.py raw.csv out.csv [n_tets]
"""
import csv
import sys


def main():
    raw, out = sys.argv[1], sys.argv[2]
    n_tets = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = [r for r in csv.reader(open(raw)) if r]
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    if not data:
        print("no kernel rows in", raw)
        return
    k = names.index("Kernel Name")
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [d[k] for d in data])
        for j, (n, u) in enumerate(zip(names, units)):
            if n in ("ID", "Process ID", "Process Name", "Host Name", "Kernel Name", "Context", "Stream", "Device", "CC"):
                continue
            vals = [d[j] if j < len(d) else "" for d in data]
            if any(v != "" for v in vals):
                w.writerow([n, u] + vals)
    d = data[0]
    get = lambda name: next((float(d[j]) for j, n in enumerate(names) if n == name and d[j] not in ("", "n/a")), None)  # noqa: E731
    dur = get("gpu__time_duration.sum")
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    print(f"{d[k]}: {dur} (unit {units[names.index('gpu__time_duration.sum')]}), registers {get('launch__registers_per_thread')}")
    if rd is not None and wr is not None:
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        tot = rd * scale[units[names.index("dram__bytes_read.sum")]] + wr * scale[units[names.index("dram__bytes_write.sum")]]
        print(f"  dram read+write: {tot / 1e6:.1f} MB" + (f" = {tot / n_tets:.1f} B per tet" if n_tets else ""))
    for name in ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum",
                 "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
                 "smsp__inst_executed.sum"):
        v = get(name)
        if v is None:
            continue
        extra = ""
        if n_tets:
            extra = (f"  = {32 * v / n_tets:.1f} per 32 tets" if "wavefront" in name or "conflict" in name
                     else f"  = {32 * v / n_tets:.0f} thread-instructions per tet")
        print(f"  {name}: {v:.4g}{extra}")
    for name in ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
                 "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
                 "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):
        v = get(name)
        if v is not None:
            print(f"  {name}: {v:.1f} %")


if __name__ == "__main__":
    main()
