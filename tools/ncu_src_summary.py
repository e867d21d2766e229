"""Summarise an `ncu --page source --csv` dump: top stalled SASS instructions and stall-reason totals.  usage: ncu_src_summary.py file.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hdr_i + 1:]:          # the first captured launch only (a second launch repeats the header)
    if r and r[0] == "Address":
        break
    if len(r) == len(hdr):
        data.append(r)
f = lambda r, h: float(r[col[h]] or 0)
tot_samples = sum(f(r, "# Samples") for r in data)
tot_inst = sum(f(r, "Instructions Executed") for r in data)
print(f"instructions executed {tot_inst:.0f}, samples {tot_samples:.0f}")
reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = {h: sum(f(r, h) for r in data) for h in reasons}
print("stall totals:", ", ".join(f"{h[6:]} {100 * v / tot_samples:.1f}%" for h, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v > 0.005 * tot_samples))
print("top instructions by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    main = max(reasons, key=lambda h: f(r, h))
    print(f"  {r[col['Address']][-6:]}  {100 * f(r, '# Samples') / tot_samples:5.2f}%  exec {f(r, 'Instructions Executed'):10.0f}  {main[6:]:14s} {r[col['Source']][:90]}")
