"""Selected metrics of an ncu report: python tools/ncu_summary.py X.ncu-rep out.csv"""
import csv, subprocess, sys
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__registers_per_thread', 'launch__shared_mem_per_block_static', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'sass__inst_executed_local_loads',
        'sass__inst_executed_local_stores', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__cycles_active.avg',
        'sm__icc_request_hit_rate.pct', 'sm__inst_issued.avg.per_cycle_active']
with open(sys.argv[2], 'w') as f:
    f.write('metric,unit,value\n')
    for h, u, v in zip(hdr, units, vals):
        if h in want or (h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued')):
            f.write('%s,%s,"%s"\n' % (h, u, v))
print(open(sys.argv[2]).read())
