"""Print the key metrics of every kernel in an ncu report (.ncu-rep): ncu_summary.py <file.ncu-rep>"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'l1tex__t_bytes.sum', 'launch__shared_mem_per_block_dynamic']
stall = [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio')]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('---', d['Kernel Name'][:70])
    for w in want:
        if w in d:
            print(f"  {w:70s} {d[w]:>16s} {units[hdr.index(w)]}")
    st = sorted(((float(d[h].replace(',', '')), h) for h in stall if d[h]), reverse=True)[:6]
    print('  top stalls:', ', '.join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st))
