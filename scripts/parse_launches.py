"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one draft+verify step."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]; kn = h.index('Kernel Name'); mv = h.index('Metric Value')
data = [(r[kn], float(r[mv].replace(',', ''))) for r in rows[hi + 1:] if len(r) > mv]
def short(n):
    n = re.sub(r'dfl::', '', n); n = re.sub(r'void ', '', n); return re.sub(r'\(.*', '', n)
lm = [i for i, d in enumerate(data) if re.search(r'gemm_skinny_kernel<\d+, 1>', d[0])]  # lm_head GEMM (argmax mode)
pairs = [(x + 1, y + 1) for x, y in zip(lm, lm[1:]) if y - x >= 20]  # (back-to-back lm_head launches = the roofline timing loop)
if not pairs:
    sys.exit(f"no complete step between two lm_head GEMMs among the {len(data)} captured launches (raise ncu's -c / -s)")
a, b = min(pairs, key=lambda p: p[1] - p[0])   # kernels after one lm_head GEMM up to and including the next = one step (no request reset inside)
step = data[a:b]
agg = collections.OrderedDict()
for n, t in step:
    k = short(n); agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += t / 1000
tot = sum(v[1] for v in agg.values())
print(f"one step: {len(step)} launches, sum of isolated durations {tot:.1f} us (cold-cache, serialised)")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:45s} x{c:3d}  {t:8.1f} us  {100 * t / tot:5.1f}%   avg {t / c:7.2f} us")
if '-v' in sys.argv:
    for n, t in step: print(f"    {short(n):45s} {t/1000:8.2f}")
