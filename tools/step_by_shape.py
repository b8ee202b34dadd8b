"""Aggregate gpurun_out/launches_step.csv (one bench step under ncu) by layer shape:
    python tools/step_by_shape.py [csv] [step_traffic.json]      (the JSON is what bench.py's roofline.traffic reads)"""
import collections, csv, importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sh = importlib.import_module("quantization---diffusion-models_b200.shapes")
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "launches_step.csv")
rows = list(csv.reader(open(path)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
dur = [float(r[ix["Metric Value"]]) for r in rows[hi + 1:] if len(r) == len(hdr) and r[ix["Metric Name"]] == "gpu__time_duration.sum"]
order = []
LAYERS = sh.sd15_unet_linears() if os.environ.get("UNFUSED") else sh.sd15_unet_linears_fused()   # bench.py's launch order
for e in LAYERS:
    order += [(e[1], e[2], e[3])] * e[4]
NL = len(order)
d = dur[-NL:]
agg = collections.OrderedDict()
for s, t in zip(order, d):
    a = agg.setdefault(s, [0, 0.0]); a[0] += 1; a[1] += t
print(f"launches {len(dur)}; sum of the step's {NL} launch durations: {sum(d) / 1e3:.1f} us")
for (m, n, k), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    g = sh.group_for(k); fl = 2 * m * n * k; by = sh.gemm_bytes_w4a16(m, n, k, g)
    ideal = max(fl / 1410.2e12, by / 6455.9e9) * 1e6
    print(f"{a[1] / 1e3:8.1f} us {a[0]:3d} x ({m:6d},{n:5d},{k:5d}) avg {a[1] / a[0] / 1e3:6.1f} us  roofline {ideal:6.1f} us  {fl / (a[1] / a[0] * 1e-9) / 1e12:7.1f} TFLOP/s")

if len(sys.argv) > 2:
    body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
    def col(metric):
        return [float(r[ix["Metric Value"]]) for r in body if r[ix["Metric Name"]] == metric][-NL:]
    rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
    unit = {r[ix["Metric Name"]]: r[ix["Metric Unit"]] for r in body}
    mul = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd_b, wr_b = sum(rd) * mul[unit["dram__bytes_read.sum"]], sum(wr) * mul[unit["dram__bytes_write.sum"]]
    names = [r[ix["Kernel Name"]] for r in body if r[ix["Metric Name"]] == "gpu__time_duration.sum"][-NL:]
    alg = sum(e[4] * sh.gemm_bytes_w4a16(e[1], e[2], e[3], sh.group_for(e[3])) for e in LAYERS)
    json.dump({"source": os.path.basename(path) + " (ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                         f"--clock-control none, one eager bench step, {NL} launches)",
               "launches": NL,
               "kernels": dict(collections.Counter(n.split("(")[0] for n in names)),
               "dram_bytes_read_per_step": rd_b, "dram_bytes_write_per_step": wr_b, "dram_bytes_per_step": rd_b + wr_b,
               "algorithmic_bytes_per_step": float(alg), "sum_launch_durations_us": sum(d) / 1e3}, open(sys.argv[2], "w"), indent=1)
