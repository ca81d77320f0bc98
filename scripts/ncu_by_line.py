import csv,re,sys,collections
linesfile, csvf, srcfile, ntile = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
addr2line={}
cur=None
for l in open(linesfile):
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*/\*([0-9a-f]{4,})\*/',l)
    if m: addr2line[int(m.group(1),16)]=cur
rows=list(csv.reader(open(csvf)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[]
for r in rows[2:]:
    if r and r[0]=="Kernel Name": break
    if len(r)>ix["stall_wait"] and r[0].startswith("0x"): data.append(r)
base=int(data[0][0],16)
agg=collections.defaultdict(lambda:[0,0])
I=lambda r,k:int(r[ix[k]] or 0)
for r in data:
    a=int(r[0],16)-base
    ln=addr2line.get(a)
    agg[ln][0]+=I(r,'Instructions Executed'); agg[ln][1]+=I(r,'# Samples')
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print('instr/tile',tot/ntile,'samples',ts)
src={}
for k in agg:
    if k and k[0] not in src:
        try: src[k[0]]=open('/root/repo/acvd_b200/csrc/'+k[0]).read().split('\n')
        except Exception: src[k[0]]=[]
for k,v in sorted(agg.items(), key=lambda kv:(kv[0] or ('',0))):
    if v[0]/ntile>=1.5 or 100*v[1]/ts>=0.8:
        text=src.get(k[0],[])[k[1]-1].strip()[:100] if k and src.get(k[0]) and k[1]-1<len(src[k[0]]) else ''
        print(f"{(k[0][:14]+':'+str(k[1])) if k else '?':>20} {v[0]/ntile:6.1f} i/t {100*v[1]/ts:5.1f}%  {text}")
