"""profiles/chol_kernel_traffic.json from an `ncu --set full` capture of agp_chol_kernel at the bench workload
(tools/profile_step.sh): DRAM bytes per launch for bench.py's roofline.traffic, tied to the kernel sources by md5 so
that a stale capture is never reported (bench.py then says so and reports null)."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
rd = float(d["dram__bytes_read.sum"][0]) * scale[d["dram__bytes_read.sum"][1]]
wr = float(d["dram__bytes_write.sum"][0]) * scale[d["dram__bytes_write.sum"][1]]
js = {"n": 2048, "particles": 64, "kernel": "agp_chol_kernel", "dram_bytes_read": rd, "dram_bytes_write": wr,
      "dram_bytes_per_launch": rd + wr, "kernel_source_md5": bench.kernel_source_md5(),
      "source": f"ncu --set full --clock-control none, one launch ({d['gpu__time_duration.sum'][0]} {d['gpu__time_duration.sum'][1]}); "
                "summary in profiles/r02_ncu_chol.txt; algorithmic bytes 2.15 GB (lower K in + lower L out, 64 particles)"}
json.dump(js, open(os.path.join(ROOT, "profiles", "chol_kernel_traffic.json"), "w"), indent=1)
print(js)
