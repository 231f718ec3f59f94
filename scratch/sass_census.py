"""SASS census of the built library -> profiles/r02_sass_census.md (run after build.sh; needs cuobjdump)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'cimrgp_b200', 'libcimrgp.so')
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(['cu++filt', n], capture_output=True, text=True).stdout.strip() or n
cols = ['DFMA', 'DADD', 'DMUL', 'MUFU', 'UBLKCP', 'STAS', 'SYNCS', 'UCGABAR', 'BAR', 'SHFL', 'WARPSYNC', 'REDUX', 'LDG', 'STG', 'LDS', 'STS', 'LDL', 'STL']
rows = []; cur = None; cnt = None
for l in sass.split('\n'):
    m = re.search(r'Function : (\S+)', l)
    if m:
        if cur: rows.append((cur, cnt))
        cur, cnt = m.group(1), collections.Counter()
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', l)
    if m and cur:
        op = m.group(1)
        cnt['total'] += 1
        for c in cols:
            if op == c or op.startswith(c + '_') or (c == 'UCGABAR' and op.startswith('UCGABAR')): cnt[c] += 1
if cur: rows.append((cur, cnt))
def short(n):
    d = demangle(n)
    d = re.sub(r'\(anonymous namespace\)::|mrgp::|void ', '', d)
    d = d.replace('(int)', '').replace('(bool)', '')
    depth = 0
    for k in range(len(d) - 1, -1, -1):      # drop the parameter list: the last parenthesised group at depth 0
        if d[k] == ')': depth += 1
        elif d[k] == '(':
            depth -= 1
            if depth == 0:
                d = d[:k]
                break
    return d.replace('<unnamed>::', '')
out = ['# SASS census of cimrgp_b200/libcimrgp.so (sm_100a), `cuobjdump -sass`, round 2 (final build)', '',
       'Static instruction counts per kernel (scratch/sass_census.py). DFMA / DADD / DMUL: FP64 pipe; UBLKCP: TMA bulk copies',
       '(`cp.async.bulk`); STAS: `st.async` (remote shared-memory stores counted on an mbarrier); SYNCS: mbarrier operations; UCGABAR: hardware',
       'cluster barrier; WARPSYNC: shuffles that the compiler had to bracket with a warp barrier (0 in the fused sweep since the warp',
       'index is a broadcast); REDUX: warp reduce; LDL / STL: local memory (spills). There are no tensor-core instructions (UTCMMA /',
       'HMMA / DMMA): the path is FP64 streaming with dy = 2 plus latency-bound small-matrix work; tensor cores do not apply (SURVEY.md §8d).', '',
       '| kernel | total | ' + ' | '.join(cols) + ' |', '|---|---|' + '---|' * len(cols)]
for n, c in sorted(rows, key=lambda r: -r[1]['total']):
    out.append('| `%s` | %d | %s |' % (short(n), c['total'], ' | '.join(str(c[k]) for k in cols)))
tot = collections.Counter()
for _, c in rows: tot.update(c)
out += ['', 'Whole library: %d kernels, %d instructions; %s.' % (len(rows), tot['total'], ', '.join('%s %d' % (k, tot[k]) for k in cols))]
open(os.path.join(ROOT, 'profiles', 'r02_sass_census.md'), 'w').write('\n'.join(out) + '\n')
print(len(rows), 'kernels', tot['total'], 'instructions')
