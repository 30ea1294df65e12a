"""Join an ncu --csv launch list (scripts/ncu_ops.py run under `ncu --profile-from-start off`) with the op list the target
wrote, and print / save the per-family roofline summary.

    python scripts/ncu_join.py conv gpurun_out/r2_ncu_conv.csv gpurun_out/r2_ops_conv.json profiles/r2_conv_traffic
    python scripts/ncu_join.py hbm  gpurun_out/r2_ncu_hbm.csv  gpurun_out/r2_ops_hbm.json  profiles/r2_hbm_kernels
"""
import csv
import io
import json
import sys
from collections import OrderedDict

what, csv_path, ops_path, out_base = sys.argv[1:5]
HBM_PEAK = 6454.0
try:
    HBM_PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
except Exception:
    pass
ops = json.load(open(ops_path))
text = open(csv_path, errors="replace").read()
start = text.index('"ID"')
rows = list(csv.DictReader(io.StringIO(text[start:])))
launches = OrderedDict()
for r in rows:
    if not r.get("ID", "").isdigit():
        continue
    d = launches.setdefault(int(r["ID"]), dict(kernel=r["Kernel Name"]))
    v = r["Metric Value"].replace(",", "")
    try:
        v = float(v)
    except ValueError:
        continue
    unit = r.get("Metric Unit", "")
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9,
             "second": 1.0, "ns": 1e-9, "us": 1e-6, "ms": 1e-3}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
L = list(launches.values())
assert len(L) == len(ops), f"{len(L)} profiled launches vs {len(ops)} ops"
fam = OrderedDict()
tot = dict(n=0, t=0.0, dram=0.0, xbar=0.0, alg=0.0, flops=0.0, tpipe=0.0)
for op, l in zip(ops, L):
    t = l.get("gpu__time_duration.sum", 0.0)
    dram = l.get("dram__bytes_read.sum", 0.0) + l.get("dram__bytes_write.sum", 0.0)
    xbar = l.get("l1tex__m_xbar2l1tex_read_bytes.sum", 0.0)
    tp = l.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
    name = l["kernel"].split("(")[0].replace("<unnamed>::", "")
    key = f"{op['tag']} {name}" if what == "conv" else name
    if what == "conv" and op["ints"]:
        # integer args of the C-ABI call: C / dilation classes separate halo- and box-mode launches
        key += " " + str(tuple(op["ints"][:8]))
    f = fam.setdefault(key, dict(n=0, t=0.0, dram=0.0, xbar=0.0, alg=0.0, flops=0.0, tpipe=0.0))
    for acc in (f, tot):
        acc["n"] += 1; acc["t"] += t; acc["dram"] += dram; acc["xbar"] += xbar; acc["alg"] += op["bytes"] or 0.0
        acc["flops"] += op["flops"] or 0.0; acc["tpipe"] += tp * t
lines = []
if what == "conv":
    lines.append(f"# {tot['n']} tagged 3x3 convolution launches of one config-2 step (batch 16, bf16), ncu per launch (cold-cache, serialised)")
    lines.append(f"{'family (tag kernel args)':100s} {'n':>4s} {'us/launch':>10s} {'TF/s':>7s} {'tensor%':>8s} {'DRAM MB':>8s} {'alg MB':>7s} {'L2->SM MB':>10s}")
    for k, f in fam.items():
        lines.append(f"{k[:100]:100s} {f['n']:4d} {f['t'] / f['n'] * 1e6:10.1f} {f['flops'] / f['t'] / 1e12:7.0f} {f['tpipe'] / f['t']:8.1f} "
                     f"{f['dram'] / f['n'] / 1e6:8.1f} {f['alg'] / f['n'] / 1e6:7.1f} {f['xbar'] / f['n'] / 1e6:10.1f}")
    summary = dict(source=f"{csv_path}: ncu per-launch counters of the {tot['n']} tagged 3x3 convolution launches of one config-2 step "
                          "(scripts/ncu_ops.py conv), this round's code",
                   launches=tot["n"], mean_dram_bytes_per_launch=tot["dram"] / tot["n"],
                   algorithmic_bytes_per_launch=tot["alg"] / tot["n"], mean_l2_to_sm_bytes_per_launch=tot["xbar"] / tot["n"],
                   mean_tensor_pipe_active_pct=tot["tpipe"] / tot["t"], summed_duration_ms=tot["t"] * 1e3,
                   tflops_under_ncu=tot["flops"] / tot["t"] / 1e12)
    lines.append(f"# all: {tot['t'] * 1e3:.3f} ms, {summary['tflops_under_ncu']:.0f} TFLOP/s under ncu, DRAM {summary['mean_dram_bytes_per_launch'] / 1e6:.1f} MB/launch "
                 f"vs algorithmic {summary['algorithmic_bytes_per_launch'] / 1e6:.1f} MB, L2->SM {summary['mean_l2_to_sm_bytes_per_launch'] / 1e6:.1f} MB/launch, "
                 f"tensor pipe {summary['mean_tensor_pipe_active_pct']:.1f} %")
    json.dump(summary, open(out_base + ".json", "w"), indent=1)
else:
    lines.append(f"# bandwidth-bound launches of one config-2 step (batch 16, bf16): ncu per launch; achieved = algorithmic bytes / duration, peak {HBM_PEAK:.0f} GB/s (measured copy)")
    lines.append(f"{'kernel':60s} {'n':>4s} {'ms total':>9s} {'alg GB/s':>9s} {'frac':>6s} {'DRAM GB/s':>10s} {'DRAM/alg':>9s}")
    for k, f in sorted(fam.items(), key=lambda kv: -kv[1]["t"]):
        lines.append(f"{k[:60]:60s} {f['n']:4d} {f['t'] * 1e3:9.3f} {f['alg'] / f['t'] / 1e9:9.0f} {f['alg'] / f['t'] / 1e9 / HBM_PEAK:6.2f} "
                     f"{f['dram'] / f['t'] / 1e9:10.0f} {f['dram'] / max(f['alg'], 1):9.2f}")
    lines.append(f"# all: {tot['n']} launches, {tot['t'] * 1e3:.3f} ms, {tot['alg'] / tot['t'] / 1e9:.0f} GB/s algorithmic")
open(out_base + ".txt", "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
