import csv,collections,sys
def launches(path):
    rows=list(csv.reader(open(path)))
    hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    agg=collections.OrderedDict()
    for r in rows[hdr+1:]:
        name=r[4].split('(')[0]; t=float(r[-1])
        agg.setdefault(name,[0,0]); agg[name][0]+=t; agg[name][1]+=1
    tot=sum(v[0] for v in agg.values())
    out=[]
    for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0]): out.append(f"{k:60s} n={v[1]:3d} total={v[0]/1e3:9.1f} us  avg={v[0]/v[1]/1e3:8.1f} us  {100*v[0]/tot:5.1f}%")
    return "\n".join(out)
if __name__=="__main__": print(launches(sys.argv[1]))
