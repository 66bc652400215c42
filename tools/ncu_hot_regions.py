import csv, subprocess, sys, collections
rep, which, lo_ex = sys.argv[1], sys.argv[2], int(sys.argv[3])
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
    # contiguous regions by ex count band
    i = 0
    while i < len(data):
        ex = int(data[i][H['Instructions Executed']])
        if ex < lo_ex: i += 1; continue
        j = i; ops = collections.Counter(); tot = 0
        while j < len(data) and int(data[j][H['Instructions Executed']]) >= lo_ex:
            src = data[j][H['Source']].strip().split()
            op = (src[1] if src[0].startswith('@') else src[0]).split('.')[0]
            ops[op] += 1; tot += int(data[j][H['Instructions Executed']]); j += 1
        print(f"rows {i}-{j} n={j-i} ex_sum={tot} first_ex={ex} :: " + ", ".join(f"{k}:{v}" for k, v in ops.most_common(12)))
        i = j
    break
