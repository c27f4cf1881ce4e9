import csv,sys
rows=[]
with open(sys.argv[1]) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
for x in r:
    if x['Metric Name']!='gpu__time_duration.sum': continue
    rows.append((x['Kernel Name'].split('(')[0], x['Grid Size'], x['Block Size'], float(x['Metric Value'].replace(',',''))))
seqs=[];cur=None
for k in rows:
    if 'msm_digits' in k[0]:
        cur=[];seqs.append(cur)
    if cur is not None: cur.append(k)
for i in [int(a) for a in sys.argv[2:]]:
    s=seqs[i]
    tot=0
    for k in s:
        if 'msm_' not in k[0] or 'precompute' in k[0]: break
        print(f"  {k[0]:40s} {k[1]:>16s} {k[2]:>12s} {k[3]/1e3:8.1f} us"); tot+=k[3]
    print(" total", tot/1e3)
