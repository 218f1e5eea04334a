"""DRAM traffic of the kernels of one policy step from `ncu --set full` captures -> profiles/r2_traffic.json.

    python profiles/traffic_from_ncu.py workload envs out.json rep1.ncu-rep [rep2.ncu-rep ...]

Each report holds one launch (captured with `-k regex:<kernel> -s <skip> -c 1` on profiles/profile_step.py at the bench
batch size); dram__bytes_read.sum + dram__bytes_write.sum of every report are added up.  bench.py reports the sum as
roofline.traffic only when workload and batch match exactly - it never scales a capture."""
import csv, io, json, subprocess, sys

workload, envs, out = sys.argv[1], int(sys.argv[2]), sys.argv[3]
kernels, total = [], 0.0
for rep in sys.argv[4:]:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    H, units, vals = rows[0], rows[1], rows[2]
    get = lambda name: (float(vals[H.index(name)].replace(",", "")), units[H.index(name)])
    def to_bytes(v, u):
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rd, wr = to_bytes(*get("dram__bytes_read.sum")), to_bytes(*get("dram__bytes_write.sum"))
    dur, du = get("gpu__time_duration.sum")
    kernels.append({"kernel": vals[H.index("Kernel Name")], "dram_bytes_read": rd, "dram_bytes_write": wr,
                    "duration_ms_under_ncu": dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "usecond": 1e-3, "msecond": 1.0, "nsecond": 1e-6}[du],
                    "report": rep})
    total += rd + wr
try:
    doc = json.load(open(out))
except Exception:
    doc = {"captures": []}
doc["captures"] = [c for c in doc["captures"] if not (c["workload"] == workload and c["envs"] == envs)]
doc["captures"].append({"workload": workload, "envs": envs, "dram_bytes": total, "kernels": kernels,
                        "source": "ncu --set full --clock-control none, one launch per kernel at step ~61 of the episode "
                                  "(profiles/profile_step.py %s %d 64)" % (workload, envs)})
json.dump(doc, open(out, "w"), indent=1)
print(json.dumps(doc["captures"][-1], indent=1))
