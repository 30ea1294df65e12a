"""Top stall lines (SASS) of the first kernel in an .ncu-rep source page."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# first kernel block
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
body = []
for r in rows[hdr_i + 1:]:
    if not r or r[0] == "Kernel Name": break
    body.append(r)
ci = hdr.index("# Samples"); si = hdr.index("Source"); ie = hdr.index("Instructions Executed")
tot = sum(int(r[ci] or 0) for r in body)
print("total samples", tot, "instructions", len(body))
ranked = sorted(enumerate(body), key=lambda t: -int(t[1][ci] or 0))[:top]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for n, r in sorted(ranked):
    st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols), reverse=True)[:2]
    print(f"{n:5d} {int(r[ci] or 0):6d} ({100*int(r[ci] or 0)/max(tot,1):4.1f}%) exec={r[ie]:>8s} {r[si].strip()[:70]:70s} {st}")
