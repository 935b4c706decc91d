"""Correlate ncu SASS-page per-instruction counts with source lines using nvdisasm -g output."""
import csv, re, sys, collections
kernel, sass_csv, dis = sys.argv[1], sys.argv[2], sys.argv[3]
# parse disassembly of the kernel: sequence of (file,line) annotations and instructions with /*addr*/
lines = open(dis).read().split('\n')
start = next(i for i,l in enumerate(lines) if l.strip().startswith('.section') and ('.text.' in l) and kernel in l)
cur=('?',0); fn=None; addr2line={}
for l in lines[start+1:]:
    if l.strip().startswith('.section'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', l)
    if m:
        cur=(m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        addr2line[int(m.group(1),16)] = cur
rows=list(csv.reader(open(sass_csv)))
h=rows[1]; ai=h.index('Address'); ii=h.index('Instructions Executed'); si=h.index('# Samples'); srci=h.index('Source')
base=None
agg=collections.Counter(); samp=collections.Counter(); tot=0; tots=0
for r in rows[2:]:
    try:
        a=int(r[ai],16) if r[ai].startswith('0x') else int(r[ai])
    except: continue
    if base is None: base=a
    off=a-base
    n=int(r[ii] or 0); s=int(r[si] or 0)
    key=addr2line.get(off,('?',0))
    agg[key]+=n; samp[key]+=s; tot+=n; tots+=s
print("total warp instr", tot, "samples", tots)
byfile=collections.Counter(); sbyfile=collections.Counter()
for k,v in agg.items(): byfile[k[0]]+=v
for k,v in samp.items(): sbyfile[k[0]]+=v
for k,v in byfile.most_common(): print(f"  {k:28s} instr {v/tot:6.3f}  samples {sbyfile[k]/max(1,tots):6.3f}")
print("top lines by samples:")
for k,v in samp.most_common(40): print(f"  {k[0]:24s}:{k[1]:5d} samples {v/max(1,tots):6.3f} instr {agg[k]/tot:6.3f}")
