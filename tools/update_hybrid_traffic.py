"""profiles/hybrid_traffic.json from an ncu launch list of the bench step with DRAM byte counters
(`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ... python bench.py --steps 2 --warmup 3
--no-also --no-cpu-baseline`, csv log): DRAM bytes of the factorisation's launches of ONE hybrid step (everything between two
Gram fills) for bench.py's roofline.traffic, tied to the kernel sources by md5 so that a stale capture is never reported."""
import csv, json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, mi, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
launches = {}
for r in rows[1:]:
    launches.setdefault(int(r[ii]), {"kernel": r[ki].split("(")[0].replace("void ", "")})[r[mi]] = float(r[vi].replace(",", ""))
ids = sorted(launches)
fills = [i for i in ids if "gramfill" in launches[i]["kernel"]]
assert len(fills) >= 2, "need two Gram fills in the capture to delimit one step"
step = [launches[i] for i in ids if fills[0] < i < fills[1]]
per_kernel = {}
for l in step:
    k = per_kernel.setdefault(l["kernel"], {"launches": 0, "dram_bytes_read": 0.0, "dram_bytes_write": 0.0, "ms": 0.0})
    k["launches"] += 1
    k["dram_bytes_read"] += l["dram__bytes_read.sum"]
    k["dram_bytes_write"] += l["dram__bytes_write.sum"]
    k["ms"] += l["gpu__time_duration.sum"] * 1e-6
rd = sum(k["dram_bytes_read"] for k in per_kernel.values())
wr = sum(k["dram_bytes_write"] for k in per_kernel.values())
js = {"n": 2048, "particles": 64, "what": "all launches of one hybrid step behind the Gram fill", "per_kernel": per_kernel,
      "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_step": rd + wr, "kernel_source_md5": bench.kernel_source_md5(),
      "source": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none over one bench step ({len(step)} launches, "
                f"{sum(k['ms'] for k in per_kernel.values()):.2f} ms serialised); algorithmic bytes 2.15 GB (lower K in + lower L out, 64 particles); "
                "the FP64 single-launch kernel moved 11.9 GB (profiles/chol_kernel_traffic.json)"}
json.dump(js, open(os.path.join(ROOT, "profiles", "hybrid_traffic.json"), "w"), indent=1)
print(json.dumps(js, indent=1))
