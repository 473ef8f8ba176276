"""ncu csv (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum per launch of one bench step) -> profiles/r02/step_traffic.json
and a per-kernel launch summary.  Usage: python scripts/make_step_traffic.py gpurun_out/launches.csv [out.json]"""
import collections, csv, json, sys
path = sys.argv[1]
out = sys.argv[2] if len(sys.argv) > 2 else "profiles/r02/step_traffic.json"
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, mi, vi, ui, gi = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
launch = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    d = launch.setdefault(r[0], {"name": r[ki], "grid": r[gi]})
    v = float(r[vi].replace(",", ""))
    u = r[ui]
    if "byte" in u.lower():
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    if r[mi].startswith("gpu__time"):
        v *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)      # -> us
    d[r[mi]] = v
L = list(launch.values())
ends = [i for i, d in enumerate(L) if "conv_post" in d["name"]]
starts = [i for i, d in enumerate(L) if "embed_kernel" in d["name"]]
full = [(a, e) for a in starts for e in ends if a < e and not any(a < x < e for x in starts + ends)]
if full:                                             # one full step inside the window: embed_kernel ... conv_post_kernel
    step = L[full[-1][0]: full[-1][1] + 1]
elif starts and ends and ends[0] < starts[0]:        # window = tail of step A + head of step B (identical steps): stitch one step
    tail, head = L[: ends[0] + 1], L[starts[0]:]
    key = [d["name"] for d in tail[:6]]
    cut = next((i for i in range(len(head)) if [d["name"] for d in head[i:i + 6]] == key and
                sum("pack_k" in d["name"] for d in head[:i]) + sum("pack_k" in d["name"] for d in tail) ==
                max(1, sum("pack_k" in d["name"] for d in tail + head) // 2 if sum("pack_k" in d["name"] for d in tail + head) > 6 else 6)), None)
    step = (head[:cut] + tail) if cut is not None else L
else:
    step = L
tot_us = sum(d.get("gpu__time_duration.sum", 0) for d in step)
tot_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in step)
tc = [d for d in step if "conv_tc_kernel" in d["name"] or "resstack" in d["name"]]
tc_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in tc)
tc_us = sum(d.get("gpu__time_duration.sum", 0) for d in tc)
by = collections.OrderedDict()
for d in step:
    n = d["name"].split("(")[0].replace("void ", "").replace("fs2::", "")
    e = by.setdefault(n, {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    e["launches"] += 1; e["us"] += d.get("gpu__time_duration.sum", 0); e["dram_bytes"] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
res = {"source": f"ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one bench step ({path}); "
                 "per-launch times are cold-cache and serialised",
       "launches_per_step": len(step), "step_us_serialised": tot_us, "step_dram_bytes": tot_b,
       "tcgen05_class_launches": len(tc), "tcgen05_class_us": tc_us, "tcgen05_class_dram_bytes": tc_b,
       "tcgen05_class_dram_bytes_per_launch": tc_b / max(len(tc), 1), "tcgen05_class_share_of_step": tc_us / max(tot_us, 1e-9),
       "by_kernel": by}
json.dump(res, open(out, "w"), indent=1)
print(json.dumps({k: v for k, v in res.items() if k != "by_kernel"}, indent=1))
for n, e in by.items():
    print(f"{n[:60]:60s} x{e['launches']:3d} {e['us']:9.1f} us {e['dram_bytes'] / 1e6:10.1f} MB")
for i, d in enumerate(step):
    print(i, d["name"].replace("void ", "").replace("fs2::", "")[:48], d["grid"], round(d.get("gpu__time_duration.sum", 0), 1), "us",
          round((d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)) / 1e6, 1), "MB")
