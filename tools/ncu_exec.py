import csv, sys, re
path = sys.argv[1]
rows = list(csv.reader(open(path)))
hdr = rows[1]; data = [r for r in rows[2:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
ii = hdr.index('Instructions Executed'); si = hdr.index('# Samples')
tot = sum(int(r[ii]) for r in data)
# group consecutive instructions with same exec count into regions
regions = []
cur = None
for i, r in enumerate(data):
    e = int(r[ii])
    if cur and cur[2] == e:
        cur[1] = i; cur[3] += e; cur[4] += int(r[si])
    else:
        cur = [i, i, e, e, int(r[si])]; regions.append(cur)
regions.sort(key=lambda x: -x[3])
print("total", tot)
for a, b, e, t, s in regions[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    ops = {}
    for r in data[a:b + 1]:
        op = r[1].strip().split()[0] if not r[1].strip().startswith('@') else r[1].strip().split()[1]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = sorted(ops.items(), key=lambda kv: -kv[1])[:6]
    print("sass %5d-%5d n=%4d exec/inst=%7d total=%8d (%4.1f%%) samples=%4d  %s" % (a, b, b - a + 1, e, t, 100.0 * t / tot, s, top))
