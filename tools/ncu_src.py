import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
si = hdr.index('# Samples'); ii = hdr.index('Instructions Executed')
tot = sum(int(r[si]) for r in data if r[si].isdigit())
toti = sum(int(r[ii]) for r in data if r[ii].isdigit())
print("total samples", tot, "total warp-insts", toti, "n sass", len(data))
idx = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:topn]
for i in sorted(idx):
    r = data[i]
    print("%5d %6s %5.1f%% exec=%8s  %s" % (i, r[si], 100.0 * int(r[si]) / max(tot, 1), r[ii], r[1].strip()[:90]))
