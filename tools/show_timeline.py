import sys
lines=open(sys.argv[1]).read().splitlines()
thr=float(sys.argv[2]) if len(sys.argv)>2 else 12
rows=[]
for ln in lines:
    p=ln.split(None,3)
    if len(p)==4 and p[2].startswith('s') and p[2][1:].isdigit():
        try: rows.append((float(p[0]),float(p[1]),int(p[2][1:]),p[3]))
        except: pass
i0=[i for i,r in enumerate(rows) if 'pack_multi' in r[3]][-1]
rows=rows[i0-3:]
t0=rows[0][0]
for r in rows:
    if r[1]>=thr:
        print(f"{r[0]-t0:8.1f} {r[1]:6.1f} s{r[2]} {r[3][:56]}")
print("end", max(r[0]+r[1] for r in rows)-t0)
