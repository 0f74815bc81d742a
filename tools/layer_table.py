import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr=None; recs=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    k=(d['ID'], d['Kernel Name'].split('(')[0])
    recs.setdefault(k,{})[d['Metric Name']]=(d['Metric Value'], d['Metric Unit'])
def num(m,k):
    v,u=m.get(k,('0',''))
    v=float(v.replace(',',''))
    return v
for (i,n),m in recs.items():
    t=num(m,'gpu__time_duration.sum')/1000
    print(i.rjust(3), n[:16].ljust(16), f'{t:8.1f} us', 'tensor%%=%5.1f' % num(m,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'),
          'dramR=%7.1f MB W=%7.1f MB' % (num(m,'dram__bytes_read.sum')/ (1e6 if m['dram__bytes_read.sum'][1]=='byte' else 1), num(m,'dram__bytes_write.sum')/(1e6 if m['dram__bytes_write.sum'][1]=='byte' else 1)), m['dram__bytes_read.sum'][1], 'lts%%=%5.1f' % num(m,'lts__throughput.avg.pct_of_peak_sustained_elapsed'))
