import csv,re,sys
from collections import Counter, defaultdict
sass=open(sys.argv[2]).read().split('\n')
tag=sys.argv[3]
# list of (lineno) per instruction in order for the kernel
start=None
for i,l in enumerate(sass):
    if '.text.' in l and tag in l: start=i;break
end=len(sass)
for i in range(start+1,len(sass)):
    if sass[i].startswith('//--------------------- .text.') and i>start+5: end=i;break
cur=None; inst=[]
for l in sass[start:end]:
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1], int(m.group(2))); continue
    m=re.search(r'^\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);',l)
    if m: inst.append((int(m.group(1),16), m.group(2).strip(), cur))
rows=list(csv.reader(open(sys.argv[1])))
# split kernels
kern=[]; 
for r in rows:
    if r and r[0]=='Kernel Name': kern.append({'name':r[1],'rows':[]}); continue
    if r and r[0]=='Address': kern[-1]['hdr']=r; continue
    if kern: kern[-1]['rows'].append(r)
k=[x for x in kern if sys.argv[4] in x['name']][0]
h={n:i for i,n in enumerate(k['hdr'])}
base=int(k['rows'][0][0],16)
byline=defaultdict(lambda: Counter())
assert len(k['rows'])==len(inst), (len(k['rows']), len(inst))
tot=Counter()
for r,(off,txt,cur) in zip(k['rows'],inst):
    op=txt.split()[0] if not txt.startswith('@') else txt.split()[1]
    d=byline[cur]
    for name,key in (('samples','# Samples'),('wf','L1 Wavefronts Shared'),('wf_ideal','L1 Wavefronts Shared Ideal'),('inst','Instructions Executed')):
        v=int(r[h[key]] or 0); d[name]+=v; tot[name]+=v
print(tot)
print("top lines by stall samples")
for cur,d in sorted(byline.items(), key=lambda x:-x[1]['samples'])[:25]: print(cur, dict(d))
print("top lines by shared wavefronts")
for cur,d in sorted(byline.items(), key=lambda x:-x[1]['wf'])[:16]: print(cur, dict(d))
