import csv,re,collections,subprocess,sys
rep=sys.argv[1]; n=int(sys.argv[2])
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fmalite.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_shared_ld.sum','smsp__inst_executed_op_shared_st.sum','lts__t_sector_hit_rate.pct','l1tex__t_sector_hit_rate.pct','smsp__cycles_active.avg','sm__cycles_elapsed.avg']
for k in range(n):
    r=rows[2+k]
    print('====', re.sub('lct::','',r[hdr.index('Kernel Name')])[:80])
    for w in want:
        if w in hdr: print(f"  {w}: {r[hdr.index(w)]} [{rows[1][hdr.index(w)]}]")
    src=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--launch-skip',str(k),'--launch-count','1'],capture_output=True,text=True).stdout
    srows=list(csv.reader(src.splitlines()))
    h=srows[1]; si=h.index('Source'); ii=h.index('Instructions Executed'); smp=h.index('# Samples')
    stall_cols=[i for i,x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x]
    agg=collections.Counter(); tot=0; st=collections.Counter(); data=[]
    for rr in srows[2:]:
        if len(rr)<=ii or not rr[ii].isdigit(): continue
        m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', rr[si]); op=m.group(2).split('.')[0] if m else '?'
        agg[op]+=int(rr[ii]); tot+=int(rr[ii]); data.append(rr)
        for i in stall_cols:
            if rr[i].isdigit(): st[h[i]]+=int(rr[i])
    print('  opcodes: '+', '.join(f"{o}:{100*v/tot:.1f}%" for o,v in agg.most_common(14)))
    ts=sum(st.values()); print('  stalls: '+', '.join(f"{o[6:]}:{100*v/ts:.1f}%" for o,v in st.most_common(10)))
    for rr in sorted(data,key=lambda r:-int(r[smp]))[:int(sys.argv[3]) if len(sys.argv)>3 else 8]:
        top=sorted(((int(rr[i]),h[i][6:]) for i in stall_cols if rr[i].isdigit() and int(rr[i])>0),reverse=True)[:2]
        print('   ',rr[smp].rjust(5), rr[si][:58].ljust(58), top)
