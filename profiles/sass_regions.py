import csv, sys, re
rows=list(csv.reader(open(sys.argv[1])))
# split sections
secs=[]; cur=None
for r in rows:
    if r and r[0]=="Kernel Name":
        cur={"name":r[1],"rows":[]}; secs.append(cur); continue
    if r and r[0]=="Address": cur["hdr"]=r; continue
    if cur is not None and r: cur["rows"].append(r)
which=int(sys.argv[2])
S=secs[which]; h={n:i for i,n in enumerate(S["hdr"])}
R=S["rows"]
tot=sum(int(r[h["# Samples"]]) for r in R)
print(S["name"], "instructions", len(R), "samples", tot)
# stall totals
for k in ["stall_barrier","stall_long_sb","stall_math","stall_mio","stall_short_sb","stall_wait","stall_selected","stall_not_selected","stall_lg","stall_branch_resolving","stall_no_inst","stall_dispatch"]:
    print("  %-24s %5.1f%%"%(k,100*sum(int(r[h[k]]) for r in R)/tot))
# regions: split at BAR.SYNC
reg=[];cur=[]
for i,r in enumerate(R):
    cur.append(i)
    if "BAR.SYNC" in r[h["Source"]] or "EXIT" in r[h["Source"]]:
        reg.append(cur);cur=[]
if cur: reg.append(cur)
out=[]
for g in reg:
    sm=sum(int(R[i][h["# Samples"]]) for i in g)
    if sm/tot<0.01: continue
    ops={}
    for i in g:
        op=R[i][h["Source"]].split()[0] if not R[i][h["Source"]].strip().startswith("@") else R[i][h["Source"]].split()[1]
        op=op.split(".")[0]
        ops[op]=ops.get(op,0)+int(R[i][h["# Samples"]])
    top=sorted(ops.items(),key=lambda x:-x[1])[:7]
    ndmma=sum(1 for i in g if "DMMA" in R[i][h["Source"]])
    execd=max(int(R[i][h["Instructions Executed"]]) for i in g)
    print("region insts %4d..%4d  samples %5.1f%%  DMMA insts %3d  maxexec %9d  top: %s"%(g[0],g[-1],100*sm/tot,ndmma,execd," ".join("%s:%.1f"%(k,100*v/tot) for k,v in top)))
