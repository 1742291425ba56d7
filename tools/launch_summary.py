"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}.get(r[ui], 1.0)
    name = r[ki].split('(')[0][:70]
    tot[name] += v
    cnt[name] += 1
T = sum(tot.values())
for k, v in tot.most_common(15):
    print('%-72s n=%5d %12.1f us  avg %9.1f us  %5.1f%%' % (k, cnt[k], v, v / cnt[k], 100 * v / T))
