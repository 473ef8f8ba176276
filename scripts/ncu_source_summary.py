"""Summarise the source page of an ncu report (csv from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`, optionally .gz):
instructions and stall samples per role of conv_tc_kernel, and the source lines with the most samples."""
import csv, gzip, sys
path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read().split("\n")
blocks, cur = [], None
for l in txt:
    if l.startswith('"Function Name"'):
        cur = {"name": l, "rows": []}; blocks.append(cur)
    elif l.startswith('"Line No"') and cur is not None:
        cur["hdr"] = next(csv.reader([l]))
    elif cur is not None and "hdr" in cur and l.startswith('"'):
        r = next(csv.reader([l]))
        if len(r) == len(cur["hdr"]) and r[0]:
            cur["rows"].append(r)
b = max(blocks, key=lambda b: len(b["rows"]))
h = b["hdr"]; ix = {}
for i, n in enumerate(h):
    ix.setdefault(n, i)
def g(r, n):
    try: return float(r[ix[n]])
    except Exception: return 0.0
rows = b["rows"]
tot_i = sum(g(r, "Instructions Executed") for r in rows); tot_s = sum(g(r, "# Samples") for r in rows)
print(b["name"][:100]); print(f"warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
# roles by source text markers (line ranges found from the banner comments)
marks = {}
for r in rows:
    src = r[1]
    for key in ("weight-stage producer", "MMA issuer", "transform warps", "epilogue warps ====", "PTX wrappers", "One 32-row x W-column block", "One K-block of one transform thread", "__global__ void"):
        if key in src: marks[key] = int(r[0])
print("markers:", marks)
stalls = [n for n in ix if n.startswith("stall_") and "Not Issued" not in n]
print(f"\ntop {top} source lines by samples:")
for r in sorted(rows, key=lambda r: -g(r, "# Samples"))[:top]:
    st = sorted(((k[6:], g(r, k)) for k in stalls), key=lambda kv: -kv[1])[:3]
    print(f"{r[0]:>5s} {100*g(r,'# Samples')/tot_s:5.1f}%  instr {g(r,'Instructions Executed')/1e6:7.2f}M  {[(k, int(v)) for k, v in st]} | {r[1].strip()[:90]}")
print("\nlines with > 0.5% of instructions:")
for r in rows:
    if g(r, "Instructions Executed") > 0.005 * tot_i:
        print(f"{r[0]:>5s} instr {g(r,'Instructions Executed')/1e6:7.2f}M ({100*g(r,'Instructions Executed')/tot_i:4.1f}%) samples {100*g(r,'# Samples')/tot_s:4.1f}% | {r[1].strip()[:100]}")
