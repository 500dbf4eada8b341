"""Print selected metrics from an `ncu --page raw --csv` dump (read here, no GPU needed)."""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
h, u = rows[0], rows[1]
pats = sys.argv[2:] or [
    r'gpu__time_duration.sum$', r'smsp__issue_active.avg.pct', r'smsp__inst_executed.sum$',
    r'smsp__inst_executed.avg.per_cycle_active$', r'sm__pipe_fma.*cycles_active.*pct_of_peak_sustained_active',
    r'sm__inst_executed_pipe_(fma|alu|fmaheavy|fmalite|lsu|xu|uniform|cbu|adu).avg.pct_of_peak_sustained_active',
    r'sm__pipe_alu_cycles_active.*pct_of_peak_sustained_active',
    r'smsp__average_warps?_issue_stalled.*_per_issue_active', r'smsp__warps_active.avg.per_cycle_active$',
    r'sm__warps_active.avg.pct_of_peak_sustained_active', r'smsp__thread_inst_executed_per_inst_executed.ratio',
    r'sm__cycles_elapsed.avg$', r'sm__cycles_active.avg$', r'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$',
    r'lts__t_bytes.sum$', r'dram__bytes_(read|write).sum$', r'launch__registers_per_thread$', r'launch__grid_size',
    r'launch__waves_per_multiprocessor', r'sm__throughput.avg.pct_of_peak_sustained_elapsed',
    r'launch__occupancy_limit', r'sm__sass_thread_inst_executed_op_f(add|mul|fma)_pred_on.sum$',
    r'smsp__sass_thread_inst_executed_op_fp32_pred_on.sum$', r'sm__inst_executed_pipe_fp32',
]
for r in rows[2:]:
    d = {h[i]: (r[i], u[i]) for i in range(len(h))}
    print('==', d.get('Kernel Name', ('?',))[0][:90])
    for k in sorted(d):
        if any(re.search(p, k) for p in pats):
            print('  %-95s %s %s' % (k, d[k][0], d[k][1]))
