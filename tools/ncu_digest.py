"""Digest of an ncu report: headline metrics, stall mix, DRAM bytes, hottest SASS lines, opcode mix.
usage: python tools/ncu_digest.py <report.ncu-rep> [n_hot_lines]"""
import csv,sys,subprocess
from collections import Counter
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','details'],capture_output=True,text=True).stdout
import re
for line in out.splitlines():
    if re.search(r"Duration|DRAM Throughput|Executed Ipc Active|Issue Slots Busy|No Eligible|Mem Pipes Busy|L1/TEX Hit|L2 Hit Rate|Registers Per|Max Bandwidth|Mem Busy",line): print(line.rstrip())
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); d=dict(zip(rows[0],rows[-1]))
ks=[k for k in d if 'smsp__average_warps_issue_stalled' in k and k.endswith('per_issue_active.ratio')]
vals=sorted(((float(d[k].replace(',','')) if d[k] not in ('','n/a') else 0),k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio','')) for k in ks)
print(' | '.join(f'{k}={v:.2f}' for v,k in vals[-9:]))
for k in ['dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__cycles_elapsed.max','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','lts__t_sectors.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']:
    print(k,d.get(k))
src=subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=rows[2:]; ix={h:i for i,h in enumerate(hdr)}
def f(r,k):
    try: return float(r[ix[k]])
    except: return 0.0
tot=sum(f(r,'# Samples') for r in data)
print('samples',tot)
for r in sorted(data,key=lambda r:-f(r,'# Samples'))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    st={k:f(r,k) for k in hdr if k.startswith('stall_') and 'Not' not in k}
    main=sorted(st.items(),key=lambda kv:-kv[1])[:2]
    print(f"{f(r,'# Samples'):7.0f} {r[ix['Source']][:64]:64s} {main}")
c=Counter(); s=Counter()
for r in data:
    t=r[ix['Source']].split()
    if not t: continue
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    c[op]+=f(r,'Instructions Executed'); s[op]+=f(r,'# Samples')
T=sum(c.values()); print('total warp instr',T)
print(' '.join(f"{op}:{v/T*100:.1f}%/{s[op]/tot*100:.1f}%" for op,v in c.most_common(14)))
