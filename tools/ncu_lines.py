import re,csv,collections,subprocess,sys
"""Per-source-line attribution of one kernel of an ncu capture:  ncu_lines.py REPORT.ncu-rep [TOP] [KERNEL-SUBSTRING] [CUBIN-BASENAME]"""
rep=sys.argv[1]
kern=sys.argv[3] if len(sys.argv)>3 else "dcb_general_kernel"
cubin=sys.argv[4] if len(sys.argv)>4 else "decombine"
subprocess.run("mkdir -p /tmp/cub && cd /tmp/cub && rm -f *.cubin && cuobjdump -xelf all /root/repo/decombinator_b200/libdcb.so >/dev/null 2>&1 && nvdisasm --print-line-info %s.sm_100a.cubin 2>&1 | awk '/\\.text\\..*%s/{f=1} /\\.text\\./{ if (f && !/%s/) exit } f' > gen_lines.txt" % (cubin, kern, kern),shell=True)
cur=None; byaddr={}
for l in open('/tmp/cub/gen_lines.txt'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',l)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m2=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m2: byaddr[int(m2.group(1),16)]=cur
raw=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=rows[2:]
base=int(data[0][ix['Address']],16)
agg=collections.Counter(); smp=collections.Counter(); thr=collections.Counter()
tot=0
for r in data:
    off=int(r[ix['Address']],16)-base
    ie=float(r[ix['Instructions Executed']]); s=int(r[ix['# Samples']]); t=float(r[ix['Thread Instructions Executed']])
    k=byaddr.get(off,('?',0))
    agg[k]+=ie; smp[k]+=s; thr[k]+=t; tot+=ie
print("total warp instrs",tot)
src={}
for k,v in agg.most_common(int(sys.argv[2]) if len(sys.argv)>2 else 30):
    try: line=open('/root/repo/decombinator_b200/csrc/'+k[0]).read().splitlines()[k[1]-1].strip()[:90]
    except Exception: line=''
    print("%5.2f%% thr %4.1f %s:%d  %s"%(100*v/tot,thr[k]/max(v,1),k[0],k[1],line))
