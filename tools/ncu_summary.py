"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.

    python tools/ncu_summary.py gpurun_out/prof_prefill.ncu-rep profiles/r01_prefill_ncu.txt
"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "sm__cycles_active.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg", "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"kernel: {d.get('Kernel Name', '?')[:160]}")
        lines.append(f"grid {d.get('Grid Size', '?')} block {d.get('Block Size', '?')}")
        for h, u, v in zip(hdr, units, r):
            if any(h.endswith(k) or h == k for k in KEYS):
                lines.append(f"  {h} = {v} {u}")
        lines.append("")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    if len(rows) > 2:
        hdr = rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
        agg = {c: sum(int(r[ix[c]] or 0) for r in data) for c in stall_cols}
        lines.append(f"warp-stall samples: {tot}")
        lines.append("  by reason: " + ", ".join(f"{c[6:]} {v}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > 0.01 * tot))
        lines.append("  top instructions (samples, SASS, dominant stall):")
        top = sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:25]
        for r in top:
            n = int(r[ix["# Samples"]] or 0)
            dom = max(stall_cols, key=lambda c: int(r[ix[c]] or 0))
            lines.append(f"    {n:7d}  {r[ix['Source']].strip()[:90]:90s} {dom[6:]}")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines[:40]))


if __name__ == "__main__":
    main()
