"""Export selected metrics of every captured launch of an ncu report to a small JSON file (for profiles/).
usage: ncu_to_json.py report.ncu-rep out.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'sass__inst_executed_register_spilling', 'sass__inst_executed_global_loads',
        'sass__inst_executed_global_stores', 'sass__inst_executed_shared_loads', 'sass__inst_executed_shared_stores',
        'smsp__warps_eligible.avg.per_cycle_active',
        'sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active',
        'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed']
want += [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and h.endswith('per_issue_active.ratio')]
res = []
for r in rows[2:]:
    d = {}
    for k in want:
        if k in hdr:
            i = hdr.index(k)
            d[k] = (r[i] + (' ' + units[i] if units[i] else '')).strip()
    res.append(d)
json.dump(res, open(out, 'w'), indent=1)
print(out, len(res), 'launches')
