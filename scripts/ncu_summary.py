"""Summarise an .ncu-rep (raw + source pages) for one kernel: key metrics, stall mix, hottest instructions."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keys = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'launch__waves_per_multiprocessor', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_op_red.sum', 'lts__t_sectors_op_atom.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed']
for r in rows[2:]:
    print("KERNEL", r[idx['Kernel Name']][:90])
    for k in keys:
        if k in idx:
            print(f"  {k:78s} {r[idx[k]]:>16s} {units[idx[k]]}")
    for h in hdr:
        if 'issue_stalled' in h and h.endswith('.ratio') and 'not_issued' not in h:
            v = float(r[idx[h]] or 0)
            if v > 0.05:
                print(f"  stall {h.split('issue_stalled_')[1].split('_per_')[0]:30s} {v:8.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[idx['# Samples']].strip().isdigit()]   # all kernels of the report
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[idx['# Samples']] or 0) for r in data)
print("total samples", tot, " instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:top_n]:
    st = {s[6:]: int(r[idx[s]] or 0) for s in stalls if int(r[idx[s]] or 0) > 0.15 * int(r[idx['# Samples']] or 1)}
    print(f"  {r[idx['Address']][-5:]} {int(r[idx['# Samples']]):6d} {r[idx['Source']][:64]:64s} {st}")
