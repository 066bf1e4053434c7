"""Prints the key ncu metrics + stall breakdown + hottest SASS regions of one kernel in a .ncu-rep.
usage: python scripts/ncu_kernel_report.py <rep> <kernel-regex> [--regions N]"""
import collections, csv, re, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
nreg = int(sys.argv[sys.argv.index("--regions") + 1]) if "--regions" in sys.argv else 12
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines())); h = rr[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__cycles_active.avg', 'sm__cycles_active.max', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum']
for r in rr[2:]:
    print("==", r[h.index('Kernel Name')][:80])
    for w in want:
        if w in h: print('  ', w, r[h.index(w)], rr[1][h.index(w)])
    st = [(x, float(r[i])) for i, x in enumerate(h) if x.startswith('smsp__pcsamp_warps_issue_stalled_') and not x.endswith('_not_issued') and r[i].replace('.', '', 1).isdigit()]
    t = sum(v for _, v in st) or 1
    print('   stalls:', ', '.join('%s %.0f%%' % (x.replace('smsp__pcsamp_warps_issue_stalled_', ''), 100 * v / t) for x, v in sorted(st, key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
secs = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {'name': r[1], 'rows': []}; secs.append(cur)
    elif r and r[0] == "Address": cur['hdr'] = r
    elif cur is not None and len(r) > 5: cur['rows'].append(r)
for s in secs:
    hh = s['hdr']; ii = hh.index("Instructions Executed"); si = hh.index("# Samples")
    stall_cols = [(i, c) for i, c in enumerate(hh) if c.startswith("stall_")]
    tot = sum(int(r[ii]) for r in s['rows']); ts = sum(int(r[si]) for r in s['rows'])
    print(s['name'][:70], 'sass', len(s['rows']), 'inst', tot, 'samples', ts)
    op = collections.Counter(); ops = collections.Counter()
    for r in s['rows']:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[1]); o = m.group(2).split('.')[0] if m else '?'
        op[o] += int(r[ii]); ops[o] += int(r[si])
    print('  opcodes:', ', '.join(f'{o} {100*c/tot:.1f}%/{100*ops[o]/ts:.1f}%' for o, c in op.most_common(14)))
    runs = []; prev = None
    for k, r in enumerate(s['rows']):
        c = int(r[ii])
        if prev is None or c != prev[0]: prev = [c, k, k, 0, 0]; runs.append(prev)
        prev[2] = k; prev[3] += c; prev[4] += int(r[si])
    runs.sort(key=lambda x: -x[4])
    for c, a, b, t, sm in runs[:nreg]:
        print(f'  rows {a}-{b} ({b-a+1} instrs) execs {c} inst {100*t/tot:.1f}% samples {100*sm/ts:.1f}%   first: {s["rows"][a][1].strip()[:50]}')
