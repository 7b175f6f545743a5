"""Per-source-line shared-memory wavefronts (total / excessive) of one kernel launch from an ncu report."""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
extra = sys.argv[3:]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"] + extra,
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2 or r[2] != "-": continue
    try:
        w = int(r[hdr.index("L1 Wavefronts Shared")]); x = int(r[hdr.index("L1 Wavefronts Shared Excessive")]); s = int(r[hdr.index("# Samples")])
    except Exception: continue
    k = (cur, int(r[0])); agg[k][0] += w; agg[k][1] += x; agg[k][2] += s; agg[k][3] = r[1][:100]
tw = sum(v[0] for v in agg.values()); tx = sum(v[1] for v in agg.values()); ts = sum(v[2] for v in agg.values())
print(f"wavefronts {tw} excessive {tx} samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]:
    print(f"{v[0]:9d} {v[1]:9d} {100*v[2]/max(ts,1):5.1f}%smp {k[0]}:{k[1]:<4d} {v[3]}")
