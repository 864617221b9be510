#!/bin/bash
# Key numbers of an ncu --set full capture (run here, on the CPU container, on a report pulled back in gpurun_out/):
#   bash scripts/ncu_extract.sh gpurun_out/r2_first/rows_group_default.ncu-rep [out.csv]
# Duration, DRAM bytes, L2 hit rate, L1 wavefronts / requests / sectors (the wavefront model of
# profiles/r1_gather_size_sweep.md: wavefronts per request = lines per instruction), occupancy, registers, and the warp
# stall reasons sorted by share.
rep=$1
raw=$(mktemp)
ncu -i "$rep" --page raw --csv > "$raw" 2>/dev/null || { echo "cannot read $rep"; exit 1; }
python - "$raw" "${2:-}" <<'PY'
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
units = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None
data = rows[2:] if units else rows[1:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_lg.sum", "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sector_hit_rate.pct",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "sm__inst_executed.sum", "smsp__inst_executed_op_global_ld.sum", "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"]
out = []
for d in data:
    rec = dict(zip(hdr, d))
    print("==", rec.get("Kernel Name", "?")[:100])
    for k in want[1:]:
        if k in rec:
            print("   %-75s %s %s" % (k, rec[k], units[hdr.index(k)] if units else ""))
    stalls = [(k, rec[k]) for k in hdr if "issue_stalled" in k and k.endswith("_per_issue_active.ratio") and rec.get(k)]
    def num(x):
        try: return float(x.replace(",", ""))
        except Exception: return 0.0
    for k, v in sorted(stalls, key=lambda kv: -num(kv[1]))[:8]:
        print("   stall %-69s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    out.append({k: rec.get(k, "") for k in want})
if sys.argv[2]:
    with open(sys.argv[2], "w", newline="") as f:
        w = csv.DictWriter(f, fieldnames=want); w.writeheader(); w.writerows(out)
PY
rm -f "$raw"
