"""Compact inventory of `ncu --set full` captures: one row per distinct kernel (its first launch) and workload.

    python tools/ncu_summary.py cfg2=raw_cfg2.csv cfg3=raw_cfg3.csv ... > profiles/r02_ncu_all_kernels.csv

Input: `ncu -i X.ncu-rep --page raw --csv` (first row metric names, second row units).  Output columns (base units):
workload, kernel, grid, block, regs, smem_dyn_bytes, us, dram_read_bytes, dram_write_bytes, dram_gbps, dram_pct,
lsu_pct, fp64_pct, warps_active_pct, launches (how many launches of that kernel the capture held).
bench.py looks a kernel's `traffic` and co-bound pipe utilisations up here by name.
"""
import csv
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3, "second": 1e6}
COLS = {
    "us": "gpu__time_duration.sum",
    "dram_read_bytes": "dram__bytes_read.sum",
    "dram_write_bytes": "dram__bytes_write.sum",
    "dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lsu_pct": "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
    "fp64_pct": "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
    "regs": "launch__registers_per_thread",
    "grid": "launch__grid_size",
    "block": "launch__block_size",
    "smem_dyn_bytes": "launch__shared_mem_per_block_dynamic",
}
OUT = ["workload", "kernel", "grid", "block", "regs", "smem_dyn_bytes", "us", "dram_read_bytes", "dram_write_bytes",
       "dram_gbps", "dram_pct", "lsu_pct", "fp64_pct", "warps_active_pct", "launches"]


def num(text, unit):
    try:
        v = float(text.replace(",", ""))
    except ValueError:
        return None
    return v * SCALE.get(unit.split("/")[0], 1.0)


def main():
    w = csv.writer(sys.stdout)
    w.writerow(OUT)
    for arg in sys.argv[1:]:
        workload, path = arg.split("=", 1)
        rows = [r for r in csv.reader(open(path, newline="")) if len(r) > 8]
        if not rows:
            continue
        hdr = next(r for r in rows if "Kernel Name" in r)
        i0 = rows.index(hdr)
        units, body = rows[i0 + 1], rows[i0 + 2:]
        ki = hdr.index("Kernel Name")
        seen = {}
        for r in body:
            name = r[ki]
            us = num(r[hdr.index(COLS["us"])], units[hdr.index(COLS["us"])]) or 0.0
            count = 1
            if name in seen:
                # keep the LONGEST launch of a kernel: predicated-off launches (N1 not needed, status != RUNNING) return at once
                count = seen[name]["launches"] + 1
                if us <= (seen[name].get("us") or 0.0):
                    seen[name]["launches"] = count
                    continue
            rec = {"workload": workload, "kernel": name, "launches": count}
            for key, metric in COLS.items():
                rec[key] = num(r[hdr.index(metric)], units[hdr.index(metric)]) if metric in hdr else None
            if rec.get("us") and rec.get("dram_read_bytes") is not None:
                rec["dram_gbps"] = (rec["dram_read_bytes"] + rec["dram_write_bytes"]) / rec["us"] / 1e3
            seen[name] = rec
        for rec in seen.values():
            w.writerow([("" if rec.get(c) is None else (f"{rec[c]:.6g}" if isinstance(rec[c], float) else rec[c])) for c in OUT])


if __name__ == "__main__":
    main()
