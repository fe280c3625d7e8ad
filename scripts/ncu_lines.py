"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.
   python scripts/ncu_lines.py <ncu-rep> <kernel-regex> [top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, "", collections.Counter()])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        line = int(r[0])
        src = r[1]
        ix = hdr.index("Instructions Executed"); sx = hdr.index("# Samples")
        try: n = int(r[ix] or 0)
        except ValueError: n = 0
        try: s = int(r[sx] or 0)
        except ValueError: s = 0
        a = agg[(cur, line)]
        a[0] += n; a[1] += s; a[2] = src.strip()[:80]
        op = r[3].split()[0] if r[3] else ""
        if op.startswith("@"): op = r[3].split()[1]
        a[3][op.split(".")[0]] += n
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp-instructions", tot, "samples", ts)
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    ops = " ".join(f"{k}:{v*100//max(a[0],1)}" for k, v in a[3].most_common(4))
    print(f"{100*a[1]/max(ts,1):5.1f}% samp {100*a[0]/max(tot,1):5.1f}% inst {f}:{l:<4} {a[2]:<80} [{ops}]")
