"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum': continue
    k = re.sub(r'\(.*', '', row['Kernel Name'])
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    tot[k] += v; cnt[k] += 1
T = sum(tot.values())
print("total %.1f us over %d launches" % (T, sum(cnt.values())))
for k, v in sorted(tot.items(), key=lambda x: -x[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]:
    print("%-58s n=%4d total=%9.1f us share=%5.1f%% avg=%8.1f us" % (k[:58], cnt[k], v, 100 * v / T, v / cnt[k]))
