"""dram bytes (read + write) per launch of every kernel in an .ncu-rep -> JSON for bench.py's roofline.traffic
    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep "workload name" > profiles/traffic_r1.json"""
import csv, io, json, subprocess, sys
rows = list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout)))
hdr, units = rows[0], rows[1]
kn, rd, wr, tm = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("gpu__time_duration.sum")
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
out = {}
for r in rows[2:]:
    name = r[kn].split("(")[0].replace("void ", "")
    b = float(r[rd]) * scale[units[rd]] + float(r[wr]) * scale[units[wr]]
    out.setdefault(name, []).append({"dram_bytes": b, "ncu_time": r[tm] + " " + units[tm]})
print(json.dumps({"workload": sys.argv[2] if len(sys.argv) > 2 else "", "source": sys.argv[1].split("/")[-1],
                  "kernels": {k: {"dram_bytes_per_launch": sum(x["dram_bytes"] for x in v) / len(v), "launches": len(v),
                                  "ncu_time": v[0]["ncu_time"]} for k, v in out.items()}}, indent=1))
