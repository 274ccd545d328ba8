import csv, sys, subprocess
rep = sys.argv[1]; per_iter = float(sys.argv[2]) if len(sys.argv)>2 else 1024*271.118
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines())); hdr=rows[0]; vals=rows[2] if len(rows)>2 else rows[1]
m=dict(zip(hdr,vals))
for k in ['gpu__time_duration.sum','sm__cycles_elapsed.max','smsp__inst_executed.sum','launch__registers_per_thread','launch__occupancy_limit_shared_mem','launch__shared_mem_per_block_dynamic','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum']:
    print(k,'=',m.get(k))
for k,v in m.items():
    if k.startswith('smsp__average_warps_issue_stalled') and k.endswith('per_issue_active.ratio') and float(v)>0.1: print(k.replace('smsp__average_warps_issue_stalled_','  stall ').replace('_per_issue_active.ratio',''),v)
print('inst per problem-iteration', float(m['smsp__inst_executed.sum'])/per_iter, ' smem wavefronts/iter', float(m['l1tex__data_pipe_lsu_wavefronts_mem_shared.sum'])/per_iter)
src = subprocess.run(['ncu','-i',rep,'--page','source','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines())); hdr=rows[1]; data=rows[2:]
iS=hdr.index('Source'); iN=hdr.index('Instructions Executed'); iSm=hdr.index('# Samples')
from collections import Counter
op=Counter(); ops=Counter(); tots=sum(int(r[iSm]) for r in data)
for r in data:
    t=r[iS].split(); o=t[1] if t[0].startswith('@') else t[0]; o='.'.join(o.split('.')[:2]) if o.startswith('IMAD') else o.split('.')[0]
    op[o]+=int(r[iN]); ops[o]+=int(r[iSm])
print(' '.join(f'{o}:{c/per_iter:.0f}({100*ops[o]/tots:.0f}%)' for o,c in op.most_common(22)))
