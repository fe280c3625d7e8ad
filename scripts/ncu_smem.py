"""Shared-memory wavefronts per source line from an ncu report (source page).
   python scripts/ncu_smem.py <ncu-rep> <kernel-regex> [top]"""
import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass",
                      "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur, hdr = None, None
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        w = hdr.index("L1 Wavefronts Shared"); wi = hdr.index("L1 Wavefronts Shared Ideal")
        ix = hdr.index("Instructions Executed")
        def num(s):
            try: return int(s or 0)
            except ValueError: return 0
        a = agg[(cur, int(r[0]))]
        a[0] += num(r[w]); a[1] += num(r[wi]); a[2] += num(r[ix]) if num(r[w]) else 0; a[3] = r[1].strip()[:75]
tot = sum(a[0] for a in agg.values()); toti = sum(a[1] for a in agg.values())
print(f"shared wavefronts {tot}  ideal {toti}  ({100.0*toti/max(tot,1):.0f}% efficient)")
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100*a[0]/max(tot,1):5.1f}%  wf {a[0]:>9} ideal {a[1]:>9} ({a[0]/max(a[1],1):.1f}x) instr {a[2]:>8}  {f}:{l:<4} {a[3]}")
