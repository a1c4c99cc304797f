"""Print per-launch time / instructions from an ncu --metrics CSV log."""
import csv, collections, sys
rows = [r for r in csv.DictReader(l for l in open(sys.argv[1]) if not l.startswith('=='))]
by = collections.OrderedDict()
for r in rows:
    key = (r['ID'], r['Kernel Name'].split('(')[0][-48:], r['Grid Size'])
    by.setdefault(key, {})[r['Metric Name']] = r['Metric Value']
for (i, n, g), m in list(by.items())[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print(f"{i:>4s} {n:50s} grid={g:14s} t={float(m.get('gpu__time_duration.sum', 0)) / 1e3:8.1f}us "
          f"inst={float(m.get('smsp__inst_executed.sum', 0)) / 1e6:7.2f}M "
          f"warps%={m.get('sm__warps_active.avg.pct_of_peak_sustained_active', '')[:5]} regs={m.get('launch__registers_per_thread')}")
