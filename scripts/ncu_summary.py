"""Summarise an .ncu-rep: headline metrics + top stall instructions (reads with `ncu -i`, no GPU needed)."""
import csv, io, subprocess, sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__pcsamp_sample_buffer_full"]
print("metric,unit,value")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w},{units[i]},{vals[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__pcsamp_warps_issue_stalled") and "not_issued" not in h:
        print(f"{h},{units[i]},{vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
idx = {k: i for i, k in enumerate(h)}
tot = sum(int(r[idx["# Samples"]]) for r in data)
print(f"# source page: {len(data)} SASS instructions, {tot} samples")
for r in sorted(data, key=lambda r: -int(r[idx["# Samples"]]))[:top_n]:
    st = {k: int(r[idx[k]]) for k in h if k.startswith("stall_") and "Not" not in k and r[idx[k]] not in ("", "-") and int(r[idx[k]]) > 0}
    st = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"# {r[idx['# Samples']]:>6} exec={r[idx['Instructions Executed']]:>8}  {r[idx['Source']].strip()[:64]:<64} {st}")
