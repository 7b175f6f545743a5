"""Per-source-line warp-stall samples of one kernel from an ncu report (needs -lineinfo + --import-source on).

  python tools/ncu_lines.py gpurun_out/x.ncu-rep regex:prior_fused [top]
"""
import csv, subprocess, sys, io, collections
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kern, "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, ""])
stall_tot = collections.Counter()
kernels_seen = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr) - 2: continue
    if r[2] != "-": continue      # SASS rows carry an address; the CUDA row has '-'
    try:
        s = int(r[hdr.index("# Samples")]); 
    except Exception: continue
    key = (cur_file, int(r[0]))
    agg[key][0] += s
    agg[key][2] = r[1][:110]
    for i, h in enumerate(hdr):
        if h.startswith("stall_") and "Not Issued" not in h:
            try: stall_tot[(key, h)] += int(r[i])
            except Exception: pass
tot = sum(v[0] for v in agg.values())
print(f"total samples {tot}")
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    st = sorted(((n, h) for (k, h), n in stall_tot.items() if k == key), reverse=True)[:3]
    print(f"{100*v[0]/max(tot,1):5.1f}%  {key[0]}:{key[1]:<4d} {v[2]}   [{', '.join(f'{h[6:]}={n}' for n,h in st)}]")
