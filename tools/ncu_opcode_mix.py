import csv, subprocess, sys, collections
rep, which = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = raw.splitlines()
blocks = []; cur = None
for ln in lines:
    if ln.startswith('"Kernel Name"'):
        cur = {"name": ln, "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(ln)
for b in blocks:
    if which not in b["name"]: continue
    rows = list(csv.reader(b["rows"])); hdr = rows[0]; H = {h: i for i, h in enumerate(hdr)}
    data = rows[1:]
    tot = sum(int(r[H['Instructions Executed']]) for r in data)
    print("total warp instrs", tot)
    # by opcode
    byop = collections.Counter()
    for r in data:
        src = r[H['Source']].strip()
        toks = src.split()
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0]
        byop[op] += int(r[H['Instructions Executed']])
    for op, c in byop.most_common(40):
        print(f"  {op:12s} {c:12d} {100*c/tot:5.1f}%")
    break
