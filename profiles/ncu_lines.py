"""Per-source-line stall samples from an .ncu-rep (needs -lineinfo + --import-source on).
usage: python profiles/ncu_lines.py rep.ncu-rep <kernel regex> [top N]"""
import csv, subprocess, sys, re, collections
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn = path = None; hdr = None
agg = collections.OrderedDict()   # (fn) -> {(file,line,src): [samples, insts, stall dict]}
for r in rows:
    if not r: continue
    if r[0] == 'File Path': path = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': fn = r[1]; continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or not re.search(pat, fn or ''): continue
    if r[0] == '': continue   # sass row
    d = dict(zip(hdr, r))
    try: s = int(d['# Samples']); n = int(d['Instructions Executed'])
    except Exception: continue
    key = (path, r[0], r[1].strip()[:90])
    e = agg.setdefault(fn, {}).setdefault(key, [0, 0, collections.Counter()])
    e[0] += s; e[1] += n
    for k, v in d.items():
        if k.startswith('stall_') and 'Not Issued' not in k and v not in ('', '0'):
            e[2][k[6:]] += int(v)
for fn, lines in agg.items():
    tot = sum(e[0] for e in lines.values()) or 1
    toti = sum(e[1] for e in lines.values()) or 1
    print(f"== {fn[:100]}  samples={tot} warp-insts={toti}")
    for (p, ln, src), e in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top]:
        st = ' '.join(f"{k}:{v}" for k, v in e[2].most_common(3))
        print(f"  {100*e[0]/tot:5.1f}%  inst {100*e[1]/toti:5.1f}%  {p}:{ln:>4s}  {src[:70]:70s} {st}")
# optional: sample share by line ranges "a-b,c-d" of the main file (4th positional arg)
if len(sys.argv) > 4:
    for fn, lines in agg.items():
        tot = sum(e[0] for e in lines.values()) or 1
        for rng in sys.argv[4].split(','):
            a, b = map(int, rng.split('-'))
            s = sum(e[0] for (p, ln, src), e in lines.items() if p.endswith('.cu') and a <= int(ln) <= b)
            st = collections.Counter()
            for (p, ln, src), e in lines.items():
                if p.endswith('.cu') and a <= int(ln) <= b: st.update(e[2])
            print(f"  lines {rng}: {100*s/tot:.1f}%  " + ' '.join(f"{k}:{v}" for k, v in st.most_common(4)))
