"""Print the figures of a bench.py JSON line that the A/B runs compare: python scripts/bench_line.py <file> [kernel substring ...]"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d.get("roofline") or {}
print(f"{d['value']:.1f} {d['unit']}  {d['ms_per_step']:.3f} ms/step  e2e {d['e2e']['value']:.1f}  conv frac {r.get('frac', 0):.4f}  conv ms {r.get('conv_ms_per_step')}")
h = d.get("hbm_kernels") or r.get("hbm_kernels") or {}
for pat in sys.argv[2:]:
    for k, v in h.items():
        if pat in k:
            print(f"  {k}: {v['launches']} launches {v['ms']:.4f} ms {v['gbs']:.0f} GB/s ({v['frac_of_hbm_peak']:.2f} of peak)")
