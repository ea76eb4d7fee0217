import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
r = list(csv.reader(raw.splitlines()))
hdr = r[0]
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__pipe_tensor_subpipe','sm__inst_executed_pipe_tensor','sm__pipe_tensor_cycles_active','lts__t_bytes.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','launch__occupancy_limit','smsp__cycles_active.avg','sm__cycles_elapsed.max','launch__waves_per_multiprocessor']
for row in r[2:]:
    print('-----')
    for h,u,v in zip(hdr, r[1], row):
        if any(h.startswith(k) or k in h for k in keys) and v not in ('','n/a'):
            print(f'  {h} [{u}] = {v}')
sass = subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass'],capture_output=True,text=True).stdout
rows = list(csv.reader(sass.splitlines()))
hdr=None; out=[]
for x in rows:
    if x and x[0]=='Address': hdr=x; continue
    if hdr and len(x)==len(hdr): out.append(dict(zip(hdr,x)))
tot=sum(int(d['# Samples'] or 0) for d in out) or 1
stalls=[k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg={k:sum(int(d[k] or 0) for d in out) for k in stalls}
print('samples',tot, {k:f'{100*v/tot:.0f}%' for k,v in sorted(agg.items(), key=lambda kv:-kv[1])[:6]})
for d in sorted(out, key=lambda d:-int(d['# Samples'] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(f"{d['# Samples']:>6} {d['Source'][:100]:100s}", {k:d[k] for k in stalls if int(d[k] or 0)>int(d['# Samples'])*0.3})
