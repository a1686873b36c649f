"""Developer tool: write the judged subset of an ncu --set full report as CSV (metric,unit,value) plus the
warp-stall sample shares and the opcode mix of the profiled kernel.

    python tools/ncu_select.py gpurun_out/prof.ncu-rep profiles/rX_ncu_full_selected.csv
"""
import collections
import csv
import re
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
names, units, vals = rows[0], rows[1], rows[-1]
want = re.compile(r'^(Kernel Name|gpu__time_duration\.sum|dram__bytes_(read|write)\.sum$|gpu__dram_throughput\.avg\.pct|'
                  r'sm__issue_active\.avg\.pct_of_peak_sustained_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|'
                  r'sm__inst_executed_pipe_(alu|fma|fmaheavy|lsu|xu|uniform|cbu|adu|tensor)[a-z_]*\.avg\.pct_of_peak_sustained_active|'
                  r'sm__warps_active\.avg\.pct_of_peak_sustained_active|smsp__inst_executed\.sum$|'
                  r'l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$|launch__(registers_per_thread$|grid_size|block_size|'
                  r'shared_mem_per_block_(static|dynamic)|occupancy_limit_[a-z_]+|waves_per_multiprocessor)|'
                  r'sm__throughput\.avg\.pct|smsp__cycles_active\.avg$|sm__cycles_elapsed\.max$)')
with open(out, 'w', newline='') as f:
    w = csv.writer(f)
    w.writerow(['metric', 'unit', 'value'])
    for n, u, v in zip(names, units, vals):
        if want.match(n):
            w.writerow([n, u, v])
    st = {n: float(v.replace(',', '')) for n, v in zip(names, vals)
          if 'pcsamp_warps_issue_stalled' in n and not n.endswith('not_issued')}
    tot = sum(st.values()) or 1.0
    for n, v in sorted(st.items(), key=lambda t: -t[1]):
        w.writerow(['stall_share.' + n.replace('smsp__pcsamp_warps_issue_stalled_', ''), '%', f'{100 * v / tot:.2f}'])
    sass = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                          capture_output=True, text=True).stdout
    srows = list(csv.reader(sass.splitlines()))
    h, data = srows[1], srows[2:]
    ix = {n: i for i, n in enumerate(h)}
    inst = collections.Counter()
    for r in data:
        m = re.match(r'(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']].strip())
        try:
            inst[m.group(2).split('.')[0] if m else '?'] += float(r[ix['Instructions Executed']])
        except ValueError:
            pass
    ti = sum(inst.values()) or 1.0
    for op, c in inst.most_common(16):
        w.writerow(['opcode_share.' + op, '% of warp instructions', f'{100 * c / ti:.2f}'])
print(open(out).read())
