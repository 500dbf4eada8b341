"""Turns the .ncu-rep captures under gpurun_out/ into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'launch__waves_per_multiprocessor', 'launch__cluster_size', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'smsp__warps_active.avg.per_cycle_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_bytes.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
STALLS = 'smsp__average_warps_issue_stalled_'
out = {}
args = sys.argv[1:]
dest = 'r01_ncu_full_summary.json'
if args and args[0].startswith('--out='):
    dest = args.pop(0)[6:]
for name in args:
    rep = os.path.join(ROOT, 'gpurun_out', name + '.ncu-rep')
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, u = rows[0], rows[1]
    launches = []
    for r in rows[2:]:
        d = {h[i]: r[i] for i in range(len(h))}
        rec = {k: (d[k] + (' ' + u[h.index(k)] if u[h.index(k)] else '')) for k in KEYS if k in d}
        rec['stalls_per_issue'] = {k[len(STALLS):-len('_per_issue_active.ratio')]: round(float(d[k]), 3) for k in d
                                   if k.startswith(STALLS) and k.endswith('_per_issue_active.ratio') and float(d[k]) >= 0.05}
        launches.append(rec)
    out[name] = launches
json.dump(out, open(os.path.join(ROOT, 'profiles', dest), 'w'), indent=1)
for k, v in out.items():
    for rec in v:
        print(k, rec['Kernel Name'][:60], rec.get('gpu__time_duration.sum'), 'fma%', rec.get('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'),
              'issue%', rec.get('smsp__issue_active.avg.pct_of_peak_sustained_active'), 'dramR', rec.get('dram__bytes_read.sum'), 'dramW', rec.get('dram__bytes_write.sum'))
