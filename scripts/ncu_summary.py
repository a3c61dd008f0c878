"""Text summary of an .ncu-rep (ncu --set full capture): per launch the duration, DRAM bytes, unit utilisation, occupancy, bank conflicts and the
top warp-stall reasons.   python scripts/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x.txt"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [('gpu__time_duration.sum', 'duration'), ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram throughput %'), ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_active', 'L1/TEX throughput %'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM throughput %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots busy %'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
        ('smsp__inst_executed.sum', 'warp instructions'), ('sm__inst_executed.avg.per_cycle_elapsed', 'IPC'), ('lts__t_sector_hit_rate.pct', 'L2 hit rate %'),
        ('l1tex__t_sector_hit_rate.pct', 'L1 hit rate %'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem bank conflicts'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum', '  of loads'), ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum', '  of stores'),
        ('l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio', 'sectors/request (global ld)'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'), ('launch__registers_per_thread', 'registers'),
        ('launch__shared_mem_per_block_dynamic', 'dyn smem'), ('launch__shared_mem_per_block_static', 'static smem'), ('launch__occupancy_limit_registers', 'occ limit regs'),
        ('launch__occupancy_limit_shared_mem', 'occ limit smem'), ('sm__maximum_warps_per_active_cycle_pct', 'theoretical occupancy %')]
print(f"# {rep}: ncu --set full --clock-control none (cold L2: ncu flushes the caches before every kernel)")
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print(f"\n== {r[col['Kernel Name']][:150]}")
    for k, name in KEYS:
        if k in col:
            print(f"  {name:32s} {r[col[k]]:>16s} {units[col[k]]}")
    stalls = []
    for h, i in col.items():
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') or (h.startswith('smsp__average_warp_latency_issue_stalled_') and h.endswith('.ratio')):
            try:
                stalls.append((float(r[i]), h.split('stalled_')[1].rsplit('.', 1)[0]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("  top stalls (warps per issue):", ', '.join(f"{n} {v:.2f}" for v, n in stalls[:6]))
