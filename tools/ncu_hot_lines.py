"""Top source lines of a kernel from `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv`.
usage: python tools/ncu_hot_lines.py file.csv [n_lines]"""
import csv, sys, collections
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 60
rows = list(csv.reader(open(path, newline='')))
fname = None; hdr = None; lines = []
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if r[0] == '' or hdr is None: continue          # SASS row
    d = dict(zip(hdr[4:], r[4:]))
    def g(k):
        try: return float(d.get(k, '0') or 0)
        except ValueError: return 0.0
    tot = g('# Samples'); bar = g('stall_barrier')
    lines.append(dict(file=fname, line=int(r[0]), src=r[1].strip()[:110], tot=tot, nb=tot - bar, bar=bar,
                      wait=g('stall_wait'), noinst=g('stall_no_inst'), lsb=g('stall_long_sb'), ssb=g('stall_short_sb'),
                      br=g('stall_branch_resolving'), sel=g('stall_selected'), inst=g('Instructions Executed'),
                      thr=g('Avg. Threads Executed')))
T = sum(l['tot'] for l in lines); NB = sum(l['nb'] for l in lines)
print('samples %d, barrier %d (%.1f%%); non-barrier %d' % (T, T - NB, 100 * (T - NB) / T, NB))
agg = collections.Counter()
for k in ('wait', 'noinst', 'lsb', 'ssb', 'br', 'sel'):
    print('  %-8s %5.1f%% of non-barrier' % (k, 100 * sum(l[k] for l in lines) / NB))
byfile = collections.Counter()
for l in lines: byfile[l['file']] += l['nb']
print('by file:', {k: '%.1f%%' % (100 * v / NB) for k, v in byfile.most_common()})
print('top lines by non-barrier samples:   %nb   noinst%  long_sb%  inst(M)  avg-threads')
for l in sorted(lines, key=lambda x: -x['nb'])[:topn]:
    print('%5.1f%%  ni %4.0f%% lsb %4.0f%%  %8.1fM thr %4.1f  %s:%d  %s' % (100 * l['nb'] / NB, 100 * l['noinst'] / max(l['nb'], 1), 100 * l['lsb'] / max(l['nb'], 1), l['inst'] / 1e6, l['thr'], l['file'], l['line'], l['src']))
