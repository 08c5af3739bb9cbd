"""profiles/traffic.json from an `ncu --set full` capture of the capture kernel inside bench.py.
usage: ncu_traffic.py <report.ncu-rep> <workload> <batch> <model>"""
import csv, io, json, os, subprocess, sys

rep, workload, batch, model = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(out)))
head, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = head.index(name)
    v, u = float(vals[i].replace(",", "")), units[i]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    return v * scale
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
dur = get("gpu__time_duration.sum")
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = os.path.join(root, "profiles", "traffic.json")
rec = {"wca_capture_attention": {"workload": workload, "batch": batch, "model": model, "dram_bytes_per_launch": rd + wr,
                                 "dram_read": rd, "dram_write": wr, "kernel": vals[head.index("Kernel Name")],
                                 "ncu_duration": dur, "ncu_duration_unit": units[head.index("gpu__time_duration.sum")],
                                 "source": os.path.basename(rep)}}
json.dump(rec, open(path, "w"), indent=1)
print(json.dumps(rec))
