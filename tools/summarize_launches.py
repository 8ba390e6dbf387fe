"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one optimizer step (between two
gather_minibatch_kernel launches) or the whole file.   python tools/summarize_launches.py file.csv [minibatch|all]"""
import collections, csv, re, sys
path = sys.argv[1]; mode = sys.argv[2] if len(sys.argv) > 2 else "minibatch"
with open(path) as f:
    lines = [l for l in f if l.startswith('"')]
rows = list(csv.DictReader(lines))
def name_of(full):
    name = re.sub(r'\(.*', '', full)
    m = re.search(r'gemm_tc_x3_kernel<\(int\)(\d+)', full)
    if m: return 'gemm_tc_x3<%s>' % m.group(1)
    if 'gemm_tc_x3_pair_kernel' in full: return 'gemm_tc_x3_pair (cta_group::2)'
    m = re.search(r'gemm_tc_kernel<\(int\)(\d+), \(bool\)(\d)', full)
    if m: return 'gemm_tc<%s,%s>' % (m.group(1), m.group(2))
    m = re.search(r'sgemm_kernel<([^>]*)>', full)
    if m: return 'sgemm<%s>' % re.sub(r'\(bool\)', '', m.group(1))
    m = re.search(r'env_step_kernel<\(bool\)(\d)', full)
    if m: return 'env_step_kernel<%s>' % m.group(1)
    return name.replace('void ', '')
def us(row):
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    return v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
if mode == "minibatch":
    idx = [i for i, r in enumerate(rows) if 'gather_minibatch' in r['Kernel Name']]
    seg = rows[idx[-2]:idx[-1]]
else:
    seg = rows
tot = collections.Counter(); cnt = collections.Counter()
for r in seg:
    n = name_of(r['Kernel Name']); tot[n] += us(r); cnt[n] += 1
T = sum(tot.values())
print("%s: %d launches, %.1f us total (ncu per-launch durations: cold cache, serialised)" % (mode, len(seg), T))
for k, v in tot.most_common(30):
    print("%-46s %9.1f us %5.1f%%  n=%-3d avg %7.1f us" % (k[:46], v, 100 * v / T, cnt[k], v / cnt[k]))
