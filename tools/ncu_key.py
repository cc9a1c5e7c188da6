"""Key metrics + stall reasons of the first kernel in an ncu report: python tools/ncu_key.py report.ncu-rep"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines())); h = r[0]; v = r[2]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'smsp__inst_executed.sum', 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_adu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'launch__registers_per_thread', 'launch__block_size', 'sm__cycles_elapsed.avg', 'smsp__warps_active.avg.per_cycle_active',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum', 'launch__shared_mem_per_block_dynamic']
for k in want:
    if k in h: print('%-75s %s' % (k, v[h.index(k)]))
st = []
for i, k in enumerate(h):
    if 'average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio'):
        try: st.append((float(v[i]), k.split('stalled_')[1].split('_per_issue')[0]))
        except Exception: pass
print('stalls/issue:', ', '.join('%s %.2f' % (n, x) for x, n in sorted(st, reverse=True)[:9]))
