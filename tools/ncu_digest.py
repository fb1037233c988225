"""Digest of the FIRST kernel in an ncu report: duration, pipes, stall mix, DRAM/L2 bytes, hottest SASS lines.
usage: python tools/ncu_digest.py <report.ncu-rep> [n_hot_lines]"""
import csv, sys, subprocess
from collections import Counter
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
d = dict(zip(rows[0], rows[2]))
def g(k):
    try: return float(d[k].replace(',', ''))
    except Exception: return float('nan')
print(d['Kernel Name'][:60], 'duration us', g('gpu__time_duration.sum'), 'regs', d.get('launch__registers_per_thread'), 'grid', d.get('launch__grid_size'), 'block', d.get('launch__block_size'))
for k in ['sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'smsp__inst_executed.sum', 'sm__inst_executed.avg.per_cycle_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__warps_active.avg.per_cycle_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
          'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'lts__t_sectors.sum', 'lts__t_sectors_srcunit_tex.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
          'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct', 'sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active']:
    if k in d: print(f'  {k} = {d[k]}')
ks = [k for k in d if 'issue_stalled' in k and k.endswith('per_issue_active.ratio')]
vals = sorted(((g(k) if d[k] not in ('', 'n/a') else 0), k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for k in ks)
print(' | '.join(f'{k}={v:.2f}' for v, k in vals[-12:]))
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and '# Samples' in r)
hdr = rows[hi]; data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r == hdr: break
    data.append(r)
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tot = sum(f(r, '# Samples') for r in data)
print('samples', tot, 'sass lines', len(data))
for r in sorted(data, key=lambda r: -f(r, '# Samples'))[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    st = {k: f(r, k) for k in hdr if k.startswith('stall_') and 'Not' not in k}
    main = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{f(r, '# Samples'):7.0f} {r[ix['Source']][:70]:70s} {main}")
c = Counter(); s = Counter(); agg = Counter()
for r in data:
    t = r[ix['Source']].split()
    if not t: continue
    op = t[1] if t[0].startswith('@') else t[0]
    op = op.split('.')[0]
    c[op] += f(r, 'Instructions Executed'); s[op] += f(r, '# Samples')
    for k in hdr:
        if k.startswith('stall_') and 'Not' not in k: agg[k] += f(r, k)
T = sum(c.values()); print('total warp instr', T)
print(' '.join(f"{op}:{v / T * 100:.1f}%/{s[op] / tot * 100:.1f}%" for op, v in c.most_common(16)))
print('stall samples:', ' '.join(f"{k[6:]}={v / tot * 100:.1f}%" for k, v in agg.most_common(10)))
