"""Summarise an ncu report of a K1 kernel: headline metrics, stall mix, hottest SASS lines of the main loop.
usage: python scripts/ncu_hot.py gpurun_out/prof.ncu-rep [min_exec_count]"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'launch__shared_mem_per_block_dynamic', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem']
for i, h in enumerate(hdr):
    if h in want or ('pcsamp_warps_issue_stalled' in h and 'not_issued' not in h and float(vals[i] or 0) > 500):
        print(f"{h}: {vals[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ie = [int(r[ix['Instructions Executed']]) for r in data]
top = collections.Counter(ie).most_common(1)[0][0] if len(sys.argv) < 3 else int(sys.argv[2])
mx = max(ie)
loop = [(i, r) for i, r in enumerate(data) if int(r[ix['Instructions Executed']]) >= mx // 2]
tot = sum(int(r[ix['# Samples']]) for _, r in loop)
print(f"main loop: {len(loop)} SASS lines, {sum(int(r[ix['Instructions Executed']]) for _, r in loop) / mx:.1f} warp-instr per iteration, {tot} samples")
st = ['stall_long_sb', 'stall_wait', 'stall_short_sb', 'stall_selected', 'stall_not_selected', 'stall_mio', 'stall_math', 'stall_dispatch',
      'stall_branch_resolving', 'stall_no_inst', 'stall_barrier']
for i, r in sorted(sorted(loop, key=lambda t: -int(t[1][ix['# Samples']]))[:40]):
    s = int(r[ix['# Samples']])
    br = ' '.join(f"{k[6:]}={r[ix[k]]}" for k in st if int(r[ix[k]]) > 0.15 * s)
    print(i, s, r[ix['Source']].strip()[:72], '|', br)
