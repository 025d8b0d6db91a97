"""Hot spots of one kernel from `ncu -i x.ncu-rep --page source --csv --print-source sass`: samples per SASS instruction
with the dominant stall reason, packed FP32 instructions summarised.  usage: python tools/ncu_hot.py sass.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
ix = {n: i for i, n in enumerate(hdr)}
stalls = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[2:]:
    if len(r) < len(hdr): continue
    try: smp = int(r[ix["# Samples"]] or 0)
    except ValueError: continue
    st = {n: int(r[ix[n]] or 0) for n in stalls}
    data.append((int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else len(data), r[ix["Source"]], smp, st, int(r[ix["Instructions Executed"]] or 0)))
tot = sum(d[2] for d in data)
packed = [d for d in data if "FMUL2" in d[1] or "FADD2" in d[1] or "FFMA2" in d[1]]
print("total samples", tot, " packed-instruction samples", sum(d[2] for d in packed), "(%.1f %%)" % (100.0 * sum(d[2] for d in packed) / tot))
agg = {}
for d in data:
    for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
print("stall totals:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
aggp = {}
for d in packed:
    for k, v in d[3].items(): aggp[k] = aggp.get(k, 0) + v
print("  on packed:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / tot) for k, v in sorted(aggp.items(), key=lambda kv: -kv[1])[:8]))
rest = sorted([d for d in data if d not in packed], key=lambda d: -d[2])[:top]
for addr, src, smp, st, ex in sorted(rest, key=lambda d: d[0]):
    dom = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print("%6x %6.2f%% exec %10d  %-60s %s" % (addr, 100.0 * smp / tot, ex, src[:60], " ".join("%s=%d" % (k[6:], v) for k, v in dom)))
