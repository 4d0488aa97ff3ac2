#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out; T=r02m
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"axis_fft_plan_kernel|stage_kernel|norm_kernel" -c 14 -f -o /tmp/${T}_axis python tools/prof_axis512.py > $O/${T}_ncu.log 2>&1; echo "rc=$?"; tail -3 $O/${T}_ncu.log
ncu -i /tmp/${T}_axis.ncu-rep --page raw --csv > $O/${T}_axis_raw.csv 2>/dev/null
python tools/ncu_summary.py cfg5s=$O/${T}_axis_raw.csv > $O/${T}_summary.csv; python - <<'PY'
import csv
for r in csv.DictReader(open("gpurun_out/r02m_summary.csv")):
    print(r["kernel"][:60], r["us"], "dramGB/s", r["dram_gbps"], "dram%", r["dram_pct"], "lsu%", r["lsu_pct"], "fp64%", r["fp64_pct"], "warps%", r["warps_active_pct"], "regs", r["regs"])
PY
python - <<'PY'
# stall breakdown and a few more metrics of the axis kernel
import csv
rows = list(csv.reader(open("gpurun_out/r02m_axis_raw.csv")))
hdr = next(r for r in rows if "Kernel Name" in r); i0 = rows.index(hdr)
ki = hdr.index("Kernel Name")
want = [h for h in hdr if ("issue_stalled" in h and "per_issue" not in h and h.endswith(".pct")) or "warp_issue_stalled" in h and h.endswith("ratio") or h in ("smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum", "smsp__inst_executed.sum")]
seen = set()
for r in rows[i0 + 2:]:
    name = r[ki][:50]
    if name in seen or ("axis" not in name and "stage_kernel<5, 6" not in name): continue
    seen.add(name)
    print("==", name)
    vals = [(h, r[hdr.index(h)]) for h in want]
    for h, v in vals:
        try:
            if float(v.replace(",", "")) != 0: print("   ", h, v)
        except ValueError: pass
PY
cp /tmp/${T}_axis.ncu-rep $O/ 2>/dev/null; ls -la $O/${T}_axis.ncu-rep
