"""summarise ncu outputs: launch list csv -> per-kernel shares; .ncu-rep raw page -> key metrics per launch"""
import csv, collections, subprocess, sys
def launches(path):
    rows=list(csv.reader(open(path)))
    for i,r in enumerate(rows):
        if 'Kernel Name' in r: hdr=r; start=i+1; break
    ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
    agg=collections.OrderedDict()
    for r in rows[start:]:
        if len(r)<=vi: continue
        name=r[ki].split('(')[0].replace('void <unnamed>::','')
        try: v=float(r[vi].replace(',',''))
        except: continue
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
    tot=sum(a[1] for a in agg.values())
    out=['%-44s %6s %12s %7s %10s'%('kernel','n','total_us','share','avg_us')]
    for k,(n,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): out.append('%-44s %6d %12.1f %6.1f%% %10.1f'%(k[:44],n,t/1e3,100*t/tot,t/n/1e3))
    return '\n'.join(out)
KEYS=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__inst_executed_op_global_red.sum','l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_red.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active']
def rep(path):
    txt=subprocess.run(['ncu','-i',path,'--page','raw','--csv'],capture_output=True,text=True).stdout
    rows=list(csv.reader(txt.splitlines())); hdr=rows[0]; units=rows[1]; out=[]
    for r in rows[2:]:
        d=dict(zip(hdr,r)); out.append('--- '+d['Kernel Name'][:90])
        for k in KEYS:
            if k in d: out.append('   %-70s %s %s'%(k,d[k],units[hdr.index(k)]))
        st=[(float(d[h]),h.replace('smsp__pcsamp_warps_issue_stalled_','')) for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued') and d[h] not in ('','n/a')]
        tot=sum(v for v,_ in st) or 1
        out.append('   stalls: '+', '.join('%s %.0f%%'%(h,100*v/tot) for v,h in sorted(st,reverse=True)[:6]))
    return '\n'.join(out)
if __name__=='__main__':
    for p in sys.argv[1:]:
        print(launches(p) if p.endswith('.csv') else rep(p))
