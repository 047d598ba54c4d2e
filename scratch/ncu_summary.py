"""Print a compact summary of an ncu report (raw page): one column per captured launch."""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__block_size',
        'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_register_spilling', 'sass__inst_executed_shared_loads',
        'sass__inst_executed_shared_stores', 'sass__inst_executed_global_loads', 'sass__inst_executed_global_stores',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed',
        'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed', 'smsp__warps_eligible.avg.per_cycle_active']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
print('kernel', [r[hdr.index('Kernel Name')][:40] for r in rows[2:]])
for k in keys:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k:75s} {units[i]:12s}", [r[i] for r in rows[2:]])
tot = {}
for h in stall:
    i = hdr.index(h)
    tot[h] = [float(r[i]) for r in rows[2:]]
for h, v in sorted(tot.items(), key=lambda kv: -kv[1][-1])[:8]:
    print(f"{h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):30s}", [f"{x:.2f}" for x in v])
