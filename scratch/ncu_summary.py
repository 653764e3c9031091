"""Summarise an ncu report of the small-d kernel: headline metrics, stall mix, cycles by code region / CUDA line."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]; iters = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 1000 / 3
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw))); hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
def g(k): return m.get(k, '?')
print("duration ms", g('gpu__time_duration.sum'), " regs", g('launch__registers_per_thread'))
print("fp64 pipe %", g('sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active'), " shared wavefronts %", g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed'),
      " issue active %", g('smsp__issue_active.avg.pct_of_peak_sustained_active'), " alu %", g('sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'))
print("bank conflicts ld/st", g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum'), g('l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_st.sum'), " wavefronts", g('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'))
inst = float(g('smsp__inst_executed.sum')); print("instructions per warp-iteration", inst / iters)
st = {k.split('issue_stalled_')[1].split('_per_')[0]: float(v) for k, v in m.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k}
print("stalls per issue:", ', '.join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
print("dram MB r/w", float(g('dram__bytes_read.sum')) / 1e6 if g('dram__bytes_read.sum') != '?' else '?', float(g('dram__bytes_write.sum')) / 1e6 if g('dram__bytes_write.sum') != '?' else '?')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
out = []; f = None
def I(x):
    try: return int(x)
    except: return 0
cyc_per_warp_iter = None
ops = collections.Counter()
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == 'File Path': f = r[1].split('/')[-1]; continue
    if len(r) > 7 and r[0].isdigit(): out.append((I(r[6]), f, int(r[0]), r[1].strip()[:90], I(r[7])))
    elif len(r) > 7 and r[0] == '' and r[3].strip() and r[3].strip() != '...':
        t = r[3].split(); op = t[1] if t[0].startswith('@') else t[0]; ops[op.split('.')[0]] += I(r[7])
T = sum(o[0] for o in out); out.sort(reverse=True)
print("top CUDA lines by samples:")
for s, f, ln, t, n in out[:int(sys.argv[3]) if len(sys.argv) > 3 else 25]: print(f"  {100*s/T:5.1f}%  {f}:{ln:<4d} n={n/iters:7.1f}/it  {t}")
print("opcode mix per warp-iteration:", ', '.join(f"{k} {v/iters:.0f}" for k, v in ops.most_common(14)))
