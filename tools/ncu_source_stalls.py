import csv, subprocess, sys
rep = sys.argv[1]
which = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
# split per kernel
blocks = []
cur = None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = {"name": ln, "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(ln)
for b in blocks:
    if which and which not in b["name"]:
        continue
    rows = list(csv.reader(b["rows"]))
    hdr = rows[0]
    H = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]
    tot = sum(int(r[H["# Samples"]]) for r in data)
    print("==", b["name"][:90], "samples", tot, "instr rows", len(data))
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = {h: sum(int(r[H[h]]) for r in data) for h in stalls}
    print("  stall totals:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.01})
    top = sorted(data, key=lambda r: -int(r[H["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]
    for r in top:
        s = int(r[H["# Samples"]])
        dom = max(stalls, key=lambda h: int(r[H[h]]))
        idx = data.index(r)
        print(f"  {idx:5d} {s:6d} {100*s/tot:5.1f}%  {dom:18s} ex={r[H['Instructions Executed']]:>8s} thr={r[H['Avg. Threads Executed']]:>5s} {r[H['Source']].strip()[:90]}")
