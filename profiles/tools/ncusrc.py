import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
iS=hdr.index("Source"); iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples"); iT=hdr.index("Avg. Threads Executed")
cols=[h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data=[]
tot_e=0
for i,r in enumerate(rows[2:]):
    try: sm=int(r[iSm]); e=int(r[iE])
    except: continue
    tot_e+=e
    st={h:int(r[hdr.index(h)] or 0) for h in cols}
    top=sorted(st.items(), key=lambda x:-x[1])[:2]
    data.append((sm,i,r[iS].strip()[:60],e,r[iT],top))
tot=sum(d[0] for d in data)
print("total samples",tot,"total inst",tot_e)
for d in sorted(data,reverse=True)[:int(sys.argv[2]) if len(sys.argv)>2 else 25]:
    print(d)
print("--- segments")
i=0
D=sorted(data,key=lambda d:d[1])
while i<len(D):
    j=i
    while j+1<len(D) and D[j+1][3]==D[i][3]: j+=1
    n=j-i+1; te=n*D[i][3]; sm=sum(d[0] for d in D[i:j+1])
    if te>400000 or sm>150: print(f"{D[i][1]:5d}-{D[j][1]:5d} n={n:4d} exec={D[i][3]:8d} thr={D[i][4]:>5} inst={te/1e6:6.2f}M samples={sm:6d} {D[i][2][:40]}")
    i=j+1
