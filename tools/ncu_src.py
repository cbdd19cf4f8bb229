"""Summarise an `ncu --page source --csv` dump (developer tool): per kernel section, the share of stall samples and
executed instructions between barriers, and the SASS lines with the most stall samples.
usage: ncu_src.py dump.csv [section index] [top lines]"""
import csv, sys
path = sys.argv[1]
sec = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = list(csv.reader(open(path)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if not starts:
    starts = [-1]
bounds = starts + [len(rows)]
print("sections:", [(k, rows[s][1][:40] if s >= 0 else "") for k, s in enumerate(starts)])
s0, s1 = bounds[sec], bounds[sec + 1]
hdr = rows[s0 + 1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[s0 + 2:s1] if len(r) == len(hdr)]
gi = lambda r, k: int(r[ix[k]] or 0)
tot = sum(gi(r, "# Samples") for r in data)
texec = sum(gi(r, "Instructions Executed") for r in data)
print("total samples", tot, "warp instr", texec)
reg = acc_s = acc_i = first = 0
for i, r in enumerate(data):
    acc_s += gi(r, "# Samples"); acc_i += gi(r, "Instructions Executed")
    if "BAR.SYNC" in r[ix["Source"]] or i == len(data) - 1:
        print(f"  region {reg} lines {first}-{i}: samples {100 * acc_s / max(tot, 1):.1f}%  instr {100 * acc_i / max(texec, 1):.1f}%")
        reg += 1; acc_s = acc_i = 0; first = i + 1
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(gi(r, s) for r in data) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(data)), key=lambda i: -gi(data[i], "# Samples"))[:top]
for i in sorted(order):
    r = data[i]
    st = {s[6:]: gi(r, s) for s in stalls if gi(r, s) > 0}
    print(f"{i:5d} {r[ix['# Samples']]:>6} {r[ix['Instructions Executed']]:>9}  {r[ix['Source']].strip()[:70]:70s} {st}")
