import csv, io, subprocess, sys
rep=sys.argv[1]
bounds=eval(sys.argv[2])  # list of (name, lo, hi) on extract.cu lines
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(out)))
hdr=None;cur="";agg={}
tot_s=tot_i=0
for r in rows:
    if len(r)==2 and r[0] in("File Path","File Name"): cur=r[1].split("/")[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)!=len(hdr) or r[0]=="": continue
    try: s=int(r[hdr.index("# Samples")]); i=int(r[hdr.index("Instructions Executed")]); ln=int(r[0])
    except ValueError: continue
    tot_s+=s; tot_i+=i
    name="other:"+cur
    if cur==(sys.argv[3] if len(sys.argv)>3 else "extract.cu"):
        for nm,lo,hi in bounds:
            if lo<=ln<=hi: name=nm;break
    a=agg.setdefault(name,[0,0]); a[0]+=s; a[1]+=i
for k,(s,i) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-28s %5.1f%% samples %5.1f%% instr  (%d warp-instr)"%(k,100*s/tot_s,100*i/tot_i,i))
