"""Brief summary of an ncu report: pipes, stalls, memory."""
import csv, subprocess, io, sys
for rep in sys.argv[1:]:
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); m = dict(zip(rows[0], rows[2]))
    keys = {'ms': 'gpu__time_duration.sum', 'regs': 'launch__registers_per_thread', 'grid': 'launch__grid_size', 'block': 'launch__block_size',
            'tensor(DMMA) pipe %': 'TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed',
            'fp64 pipe %': 'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
            'issue active %': 'smsp__issue_active.avg.pct_of_peak_sustained_active',
            'lsu wavefronts %': 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
            'shared wavefronts': 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'shared bank conflicts': 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'L1 hit %': 'l1tex__t_sector_hit_rate.pct', 'L2 hit %': 'lts__t_sector_hit_rate.pct',
            'dram read MB': 'dram__bytes_read.sum', 'dram write MB': 'dram__bytes_write.sum', 'warps active %': 'sm__warps_active.avg.pct_of_peak_sustained_active'}
    print(rep)
    for k, v in keys.items():
        print(f"  {k}: {m.get(v)}")
    st = {k.split('issue_stalled_')[1].split('_per_')[0]: float(v) for k, v in m.items() if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and 'not_issued' not in k}
    print('  stalls per issue:', ', '.join(f"{k} {v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
