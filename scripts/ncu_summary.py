"""Text summary of an .ncu-rep (ncu --set full): per captured launch the numbers the design notes quote -- duration, grid,
registers, occupancy, issue slots, pipe utilisation, cache hit rates, DRAM bytes and the top stall reasons."""
import csv, subprocess, sys

def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    pick = [
        ("gpu__time_duration.sum", "duration"), ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("launch__cluster_size", "cluster"),
        ("launch__cluster_max_active", "clusters resident at once"), ("launch__registers_per_thread", "registers/thread"),
        ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"), ("launch__occupancy_limit_registers", "CTAs/SM (register limit)"),
        ("launch__occupancy_limit_shared_mem", "CTAs/SM (smem limit)"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("smsp__inst_executed.sum", "warp instructions"), ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / instruction"),
        ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy % (active cycles)"), ("sm__cycles_active.avg", "SM active cycles (avg)"),
        ("sm__cycles_elapsed.max", "SM elapsed cycles (max)"), ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe % (avg)"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe % (avg)"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe % (avg)"), ("sm__inst_executed_pipe_xu.max.pct_of_peak_sustained_active", "XU pipe % (max SM)"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe % (avg)"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("lts__t_sectors.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM written"), ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ]
    stalls = [n for n in hdr if n.startswith("smsp__average_warps_issue_stalled_") and n.endswith("_per_issue_active.ratio")]
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        print("=" * 100)
        print(r[col["Kernel Name"]] if "Kernel Name" in col else "?", " id", r[col["ID"]] if "ID" in col else "")
        for name, label in pick:
            if name in col:
                print(f"  {label:42s} {r[col[name]]:>18s} {units[col[name]]}")
        top = sorted(((float(r[col[n]] or 0), n) for n in stalls), reverse=True)[:6]
        print("  top stalls (warps stalled per issue):", ", ".join(f"{n.split('stalled_')[1].split('_per_issue')[0]} {v:.2f}" for v, n in top))

if __name__ == "__main__":
    for p in sys.argv[1:]:
        print("#", p)
        main(p)
