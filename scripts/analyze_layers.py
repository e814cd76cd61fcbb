import sys
rows=[]
for l in open(sys.argv[1] if len(sys.argv)>1 else 'gpurun_out/layers.md'):
    if not l.startswith('| ') or l.startswith('| layer'): continue
    f=[x.strip() for x in l.strip().strip('|').split('|')]
    rows.append((f[0], float(f[1]), float(f[2]), float(f[3])))
print('total %.3f ms'%sum(r[1] for r in rows))
groups={}
for n,ms,gf,tf in rows:
    if n.startswith('backbone.layers.'):
        st=n.split('.')[2]; kind=n.split(' ')[0].split('.')[-1]
        key='stage%s.%s'%(st, kind)
    else:
        key=n.split(' ')[0].split('.')[0]+'.'+ (n.split(' ')[0].split('.')[1] if '.' in n.split(' ')[0] else '')
    g=groups.setdefault(key,[0,0.0,0.0]); g[0]+=1; g[1]+=ms; g[2]+=gf
for k,(c,ms,gf) in sorted(groups.items(), key=lambda kv:-kv[1][1])[:16]:
    print('%-28s n=%3d  %.3f ms  %.1f GF  %.0f TF/s'%(k,c,ms,gf,gf/ms if ms else 0))
print()
seen=set()
for r in sorted(rows,key=lambda r:-r[1]):
    key=r[0].split(' ',1)[1] if ' ' in r[0] else r[0]
    if key in seen: continue
    seen.add(key)
    if len(seen)>int(sys.argv[2]) if len(sys.argv)>2 else 26: break
    print('%.4f ms %7.2f GF %5.0f TF/s  %s'%(r[1],r[2],r[3],r[0]))
