"""Hot instruction footprint of a kernel: joins the per-instruction execution counts of an ncu capture with the line table of the
cubin and lists, per source function, the bytes of code that run at least `thr` times per pop.
    cuobjdump -xelf all libavp_b200.so; nvdisasm -g avp_api.sm_100a.cubin > all.sass
    ncu -i X.ncu-rep --page source --print-source sass --csv > sass.csv
    python tools/hot_footprint.py all.sass sass.csv <kernel symbol> <pops in the captured launch> [thr]"""
import csv, re, sys, collections, os
sass, page, sym, pops = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 0.5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
cur = None; off2line = {}; inside = False
for l in open(sass):
    if l.startswith('//---------------------'):
        inside = ('.text.' + sym + ' ') in l
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*?)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,6})\*/', l)
    if m: off2line[int(m.group(1), 16)] = cur
funcs = {}
def fn(f, ln):
    if f not in funcs:
        b = []
        try:
            for i, s in enumerate(open(os.path.join(ROOT, 'automatedvaletparking_b200', 'csrc', f)), 1):
                m = re.match(r'^\s*(?:template <[^>]*>\s*)?(?:__device__|__global__|AVP_HD|static).*?([A-Za-z_0-9]+)\s*\(', s)
                if m and not s.strip().startswith('//') and m.group(1) not in ('static_assert', '__launch_bounds__'): b.append((i, m.group(1)))
        except OSError: pass
        funcs[f] = b
    name = '(kernel body)'
    for i, n in funcs[f]:
        if i <= ln: name = n
        else: break
    return name
rows = list(csv.reader(open(page, newline='')))
hdr = None; ins = []
for r in rows:
    if r and r[0] == 'Address': hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    try: ins.append((int(d['Address'], 16), float(d.get('Instructions Executed', '0') or 0)))
    except ValueError: pass
base = min(a for a, _ in ins)
hot = collections.Counter(); ex = collections.Counter(); tot_b = 0
for a, e in ins:
    fl = off2line.get(a - base)
    key = (fl[0], fn(fl[0], fl[1])) if fl else ('?', '?')
    ex[key] += e
    if e / pops >= thr: hot[key] += 16; tot_b += 16
print('kernel %s: %d instructions = %.1f KB of code; executed at least %.2f times per pop: %.1f KB; warp instructions per pop %.0f'
      % (sym, len(ins), len(ins) * 16 / 1024, thr, tot_b / 1024, sum(e for _, e in ins) / pops))
for k, b in hot.most_common(40):
    print('%6.1f KB  %7.0f warp instructions per pop  %s:%s' % (b / 1024, ex[k] / pops, k[0], k[1]))
