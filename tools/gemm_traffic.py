"""Sum the DRAM traffic of every GEMM launch of one train step from an `ncu --metrics dram__bytes_read.sum,
dram__bytes_write.sum,gpu__time_duration.sum --csv` log -> JSON read by bench.py (roofline.traffic)."""
import csv, json, sys, collections
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
tot = collections.defaultdict(float); launches = set()
for row in csv.DictReader(lines):
    v = float(row["Metric Value"].replace(",", "")); u = row["Metric Unit"]; name = row["Metric Name"]
    if name.startswith("dram__bytes"):
        v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    else:
        v *= {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1e-9)
    tot[name] += v; launches.add(row["ID"])
print(json.dumps({"kernel": "gemm_split_kernel (all GEMM launches of one C3 train step)", "launches": len(launches),
                  "dram_bytes_read": tot["dram__bytes_read.sum"], "dram_bytes_write": tot["dram__bytes_write.sum"],
                  "dram_bytes": tot["dram__bytes_read.sum"] + tot["dram__bytes_write.sum"],
                  "ncu_time_s": tot["gpu__time_duration.sum"],
                  "how": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_split (cold-cache, serialised)"}))
