"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X.csv ...`) per kernel."""
import collections, csv, sys
src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(src)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    name = r[ik].split("(")[0].replace("void ", "").replace("b200pir::<unnamed>::", "")
    v = float(r[iv].replace(",", "")) * {"ms": 1000.0, "us": 1.0, "ns": 0.001}.get(r[iu], 1.0)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
SETUP = ("k_db_synth", "k_db_to_tc5", "k_db_to_frag", "k_narrow", "k_ntt32", "at::")
tot = sum(v for k, (n, v) in agg.items() if not k.startswith(SETUP))
with open(out, "w") as f:
    f.write("# %s\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES).\n"
            "Shares exclude the one-off setup kernels (database synthesis / re-tiling, parameter upload).\n\n"
            "| kernel | launches | total us | share of step kernels |\n|---|---|---|---|\n" % title)
    for k, (n, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("| %s | %d | %.1f | %s |\n" % (k, n, v, "" if k.startswith(SETUP) else "%.1f%%" % (100 * v / tot)))
print(open(out).read())
