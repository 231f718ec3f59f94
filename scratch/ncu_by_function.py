"""Stall samples of the fused sweep by source function: ncu source page (SASS rows) aligned with nvdisasm --print-line-info of the same build."""
import csv, re, collections, sys
rep_csv, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(rep_csv)))
h = rows[1]; idx = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) >= len(h)]
lines = open(sass).read().split('\n')
start = None
for i, l in enumerate(lines):
    if l.startswith('.text.') and 'k_ci_sweepILi30' in l: start = i
    elif start is not None and l.startswith('\t.section') and i > start + 5: end = i; break
else: end = len(lines)
B = [(45, 167, 'helpers (sums, barriers)'), (168, 189, 'p1_finish'), (190, 213, 'mid1_layer0'), (214, 241, 'mid1_upper'), (242, 263, 'bias_noise_update'),
     (264, 282, 's2_region'), (283, 338, 'mid2_stats_layer0'), (339, 476, 'mid2_stats_upper'), (477, 518, 'omega_eval'), (519, 669, 'omega_solve_warp'),
     (670, 712, 'bingham2_chain'), (713, 729, 'digamma_chain'), (730, 893, 'shared_step'), (894, 944, 'layer_mid1/finish/publish'), (945, 1129, 'k_ci_sweep body')]
def bucket(ln):
    for a, b, n in B:
        if a <= ln <= b: return n
    return 'other'
ctx = 'other'; per = []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m and m.group(1).endswith('chain.cu'): ctx = bucket(int(m.group(2)))
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', l): per.append(ctx)
print('instructions: ncu', len(data), 'nvdisasm', len(per))
st = [n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
agg = collections.defaultdict(lambda: collections.Counter()); ex = collections.Counter(); ninstr = collections.Counter()
for r, b in zip(data, per):
    for n in st:
        try: agg[b][n] += int(r[idx[n]])
        except ValueError: pass
    try: ex[b] += int(r[idx['Instructions Executed']])
    except ValueError: pass
    ninstr[b] += 1
print('%-28s %7s %9s %8s | top stalls (samples)' % ('function', 'instrs', 'executed', 'samples'))
for b in sorted(agg, key=lambda k: -sum(agg[k].values())):
    tot = sum(agg[b].values())
    top = ', '.join('%s %d' % (k.replace('stall_', ''), v) for k, v in agg[b].most_common(5))
    print('%-28s %7d %9d %8d | %s' % (b, ninstr[b], ex[b], tot, top))
