import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i,r in enumerate(rows) if r and r[0]=='ID')
hdr = rows[hi]; data = rows[hi+1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
seq=[]
for r in data:
    if len(r)<=vi: continue
    v=float(r[vi].replace(',',''))*{'ns':1e-3,'us':1,'ms':1e3}.get(r[ui],1)
    seq.append((re.sub(r'\(.*','',r[ki])[:90], v))
idx=[i for i,(k,v) in enumerate(seq) if 'k_msm_digits' in k]
for n,start in enumerate(idx):
    if n % 4 != 3: continue
    tot=0
    print('---- MSM run starting at launch', start)
    for k,v in seq[start:start+40]:
        if any(x in k for x in ('k_fixed','k_tb','k_precompute','k_fq_array','elementwise')): break
        if tot and 'k_msm_digits' in k: break
        print(f'{v:10.1f} us  {k}'); tot+=v
    print(f'{tot:10.1f} us  total')
