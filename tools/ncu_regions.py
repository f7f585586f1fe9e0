"""Summarise an `ncu --page source --csv` dump of one kernel: stall samples per warp role (code region)."""
import csv, collections, sys
def load(path):
    rows=list(csv.reader(open(path)))
    hdr=rows[1]
    isrc=hdr.index('Source'); isamp=hdr.index('# Samples'); iex=hdr.index('Instructions Executed')
    stall=[(i,h.replace('stall_','')) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data=[r for r in rows[2:] if len(r)>=len(hdr)-2 and r[isamp].isdigit()]
    data=data[:len(data)//2] if len(data)%2==0 and data[0][isrc]==data[len(data)//2][isrc] else data
    return hdr,data,isrc,isamp,iex,stall
def main(path, nb=40):
    hdr,data,isrc,isamp,iex,stall=load(path)
    n=len(data); tot=sum(int(r[isamp]) for r in data)
    print(path,'instrs',n,'samples',tot,'executed',sum(int(r[iex]) for r in data))
    # region boundaries: find marker instructions
    marks=[i for i,r in enumerate(data) if 'UTMALDG' in r[isrc] or 'UTCHMMA' in r[isrc] or 'UTMASTG' in r[isrc] or 'LDTM' in r[isrc] or 'FFMA2' in r[isrc]]
    kinds=[('tma',[i for i,r in enumerate(data) if 'UTMALDG' in r[isrc]]),('mma',[i for i,r in enumerate(data) if 'UTCHMMA' in r[isrc]]),
           ('epi',[i for i,r in enumerate(data) if 'LDTM' in r[isrc] or 'UTMASTG' in r[isrc]]),('dw',[i for i,r in enumerate(data) if 'FFMA2' in r[isrc]])]
    for k,v in kinds: print(k, (min(v),max(v)) if v else None)
    step=max(1,n//nb)
    for b in range(0,n,step):
        seg=data[b:b+step]; s=sum(int(r[isamp]) for r in seg); e=sum(int(r[iex]) for r in seg)
        c=collections.Counter()
        for r in seg:
            for i,h in stall:
                if i<len(r) and r[i].isdigit(): c[h]+=int(r[i])
        if s>tot*0.004: print('%5d %6d %5.1f%% exec %10d  %s'%(b,s,100.0*s/tot,e,dict(c.most_common(4))))
if __name__=='__main__': main(sys.argv[1], int(sys.argv[2]) if len(sys.argv)>2 else 40)
