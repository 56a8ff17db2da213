"""Turn an ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch into
profiles/<round>/traffic.json: average DRAM bytes and duration per launch for every kernel (bench.py reads it
to fill roofline.traffic for the dominant kernel)."""
import csv
import json
import sys
from collections import defaultdict

src, dst, tiles = sys.argv[1], sys.argv[2], int(sys.argv[3])
lines = [l for l in open(src) if not l.startswith("==")]
acc = defaultdict(lambda: defaultdict(list))
per_id = defaultdict(dict)
for r in csv.DictReader(lines):
    name = r["Kernel Name"].split("(")[0].split("<")[0].replace("void ", "").strip()
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(unit, 1)
    per_id[(name, r["ID"])][r["Metric Name"]] = v * scale
# the bench also launches 128-tile chunks (host-buffer path); keep the full-batch launches only
longest = defaultdict(float)
for (name, _), m in per_id.items():
    longest[name] = max(longest[name], m.get("gpu__time_duration.sum", 0.0))
for (name, _), m in per_id.items():
    if m.get("gpu__time_duration.sum", 0.0) >= 0.6 * longest[name]:
        for k, v in m.items():
            acc[name][k].append(v)
out = {"tiles_per_launch": tiles, "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
       "--clock-control none on `python bench.py --steps 1 --warmup 1` (1024 conic tiles per launch)", "kernels": {}}
for k, m in acc.items():
    rd = m.get("dram__bytes_read.sum", [0]); wr = m.get("dram__bytes_write.sum", [0]); t = m.get("gpu__time_duration.sum", [0])
    out["kernels"][k] = {"launches": len(rd), "dram_read_bytes": sum(rd) / len(rd), "dram_write_bytes": sum(wr) / len(wr),
                         "duration_s_under_ncu": sum(t) / len(t)}
json.dump(out, open(dst, "w"), indent=1)
print(json.dumps(out["kernels"], indent=1))
