"""Top source lines by warp-stall samples for one kernel of an ncu report (needs -lineinfo + --import-source on).
   python profiles/hotlines.py gpurun_out/prof.ncu-rep k_solve [N]"""
import csv, subprocess, sys, collections
rep, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kern], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if len(r) > 4 and "Warp Stall Sampling (All Samples)" in r)
hdr = rows[hi]
ws = hdr.index("Warp Stall Sampling (All Samples)")
ie = hdr.index("Instructions Executed")
src_cols = [i for i, h in enumerate(hdr) if h == "Source"]
agg = collections.OrderedDict()
cur = None
for r in rows[hi + 1:]:
    if len(r) != len(hdr):
        continue
    line, text = r[0], r[src_cols[0]]
    if line.strip():
        cur = (line, text.strip())
        agg.setdefault(cur, [0, 0])
    if cur is None:
        continue
    if r[ws].isdigit() and r[2].startswith("0x"):
        agg[cur][0] += int(r[ws])
        agg[cur][1] += int(r[ie]) if r[ie].isdigit() else 0
tot = sum(v[0] for v in agg.values()) or 1
print(f"# {kern}: {tot} stall samples")
for (line, text), (s, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*s/tot:5.1f}%  inst {n:9d}  L{line:>4s}  {text[:120]}")
