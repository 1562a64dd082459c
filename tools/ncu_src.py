"""Top stall instructions of one kernel in an .ncu-rep (source page).  usage: ncu_src.py <rep> <kernel-index> [topn]"""
import csv, sys, subprocess
rep, kid = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-kernel-base","function"], capture_output=True, text=True).stdout
blocks = out.split('"Kernel Name"')
blk = blocks[int(kid)+1]
lines = blk.splitlines()
rows = list(csv.reader(lines[1:]))
hdr = rows[0]
i_src, i_samp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [(h,i) for i,h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[1:]:
    if len(r) < len(hdr): continue
    try: s = int(r[i_samp])
    except: continue
    data.append((s, r))
tot = sum(s for s,_ in data)
print("kernel", lines[0][:80], "total samples", tot)
for s, r in sorted(data, key=lambda x:-x[0])[:topn]:
    st = sorted([(int(r[i] or 0), h) for h,i in stall_cols], reverse=True)[:2]
    print(f"{s:7d} {100*s/tot:5.1f}%  {r[i_src].strip()[:70]:70s} {st}")
