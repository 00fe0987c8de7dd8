"""Top stall locations of one kernel from an .ncu-rep (source page).  usage: ncu_stalls.py rep kernel_regex [n]"""
import csv, io, subprocess, sys
rep, kre = sys.argv[1], sys.argv[2]
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", f"regex:{kre}"], capture_output=True, text=True).stdout
# the csv holds one block per launch; take the first block
blocks = raw.split('"Kernel Name"')
blk = '"Kernel Name"' + blocks[1]
rows = list(csv.reader(io.StringIO(blk)))
print(rows[0][:2])
hdr = rows[1]; data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iN = hdr.index('Source'), hdr.index('# Samples')
stalls = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iN] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
for r in sorted(data, key=lambda r: -int(r[iN] or 0))[:n]:
    st = {hdr[i][6:]: int(r[i] or 0) for i in stalls if int(r[i] or 0) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{int(r[iN]):6d} {100*int(r[iN])/max(tot,1):5.1f}%  {r[iS][:80]:80s} {st}")
agg = {hdr[i][6:]: sum(int(r[i] or 0) for r in data) for i in stalls}
print(sorted(agg.items(), key=lambda kv: -kv[1])[:8])
