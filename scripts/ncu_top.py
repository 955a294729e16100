"""Print the hottest SASS lines of an ncu --page source --csv dump: python scripts/ncu_top.py file.csv [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
H = rows[1]; ci = {h: i for i, h in enumerate(H)}
stalls = [h for h in H if h.startswith("stall_") and "Not Issued" not in h]
data = []
for k, r in enumerate(rows[2:]):
    if len(r) < len(H): continue
    data.append((float(r[ci["# Samples"]] or 0), k, r))
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
agg = {s: sum(float(d[2][ci[s]] or 0) for d in data) for s in stalls}
print("stall mix:", ", ".join("%s %.1f%%" % (s[6:], 100 * v / tot) for s, v in sorted(agg.items(), key=lambda t: -t[1]) if v > 0.01 * tot))
for v, k, r in sorted(data, key=lambda t: -t[0])[:N]:
    top = sorted(((float(r[ci[s]] or 0), s[6:]) for s in stalls), reverse=True)[:2]
    print("%5d %6.0f %5.1f%%  %-70s exec=%s  %s" % (k, v, 100 * v / tot, r[ci["Source"]].strip()[:70], r[ci["Instructions Executed"]], " ".join("%s:%.0f" % (n, x) for x, n in top if x > 0)))
