#!/usr/bin/env python
"""Like ncu_lines.py but aggregates by enclosing function (PRL_D ... name( ) of the innermost source line)."""
import collections, csv, re, subprocess, sys, os
sys.path.insert(0, os.path.dirname(__file__))
from ncu_lines import line_map

def func_table(path):
    tab=[]; cur='?'
    for i,l in enumerate(open(path),1):
        m=re.match(r'\s*(?:template <[^>]*>\s*)?(?:static\s+)?(?:__device__|PRL_D|PRL_HD|__global__)[^;{]*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(',l)
        if m and not l.strip().startswith('//'): cur=m.group(1)
        tab.append(cur)
    return tab

def main():
    src,kernel,cubin,srcdir=sys.argv[1:5]
    lm=line_map(cubin,kernel)
    tabs={}
    rows=list(csv.reader(open(src)))
    hi=next(i for i,r in enumerate(rows) if '# Samples' in r); h=rows[hi]; ci={n:i for i,n in enumerate(h)}
    base=None; agg=collections.defaultdict(collections.Counter)
    for r in rows[hi+1:]:
        if len(r)<len(h): continue
        addr=int(r[0],16); base=addr if base is None else base
        f,ln=lm.get(addr-base,('?',0))
        if f not in tabs:
            pth=os.path.join(srcdir,f); tabs[f]=func_table(pth) if os.path.exists(pth) else []
        fn=tabs[f][ln-1] if 0<ln<=len(tabs[f]) else '?'
        a=agg[(f,fn)]; a['samples']+=int(r[ci['# Samples']]); a['inst']+=int(r[ci['Instructions Executed']]); a['tinst']+=int(r[ci['Thread Instructions Executed']]); a['sass']+=1
    ts=sum(a['samples'] for a in agg.values()); ti=sum(a['inst'] for a in agg.values())
    print(f"{'function':44s} {'samp%':>6s} {'inst%':>6s} {'lanes':>5s} {'sass':>5s}")
    for k,a in sorted(agg.items(), key=lambda kv:-kv[1]['inst'])[:40]:
        print(f"{k[0][:14]+':'+k[1]:44s} {100*a['samples']/ts:6.2f} {100*a['inst']/ti:6.2f} {a['tinst']/max(1,a['inst']):5.1f} {a['sass']:5d}")
if __name__ == "__main__":
    main()
