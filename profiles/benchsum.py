"""Print the per-kernel table of bench.py JSON lines.
usage: python profiles/benchsum.py a.json [b.json ...]   (or JSON lines on stdin when no file is given)"""
import json, sys
texts = [open(f).read() for f in sys.argv[1:]] if len(sys.argv) > 1 else [sys.stdin.read()]
for text in texts:
    for line in text.strip().splitlines():
        if not line.startswith('{'): continue
        d = json.loads(line)
        e2e = d.get("e2e", {}).get("value")
        print("ms/step %.3f  img/s %.1f  e2e %s  %s" % (d["ms_per_step"], d["value"], "%.1f" % e2e if e2e else "-",
                                                      d.get("config", {}).get("workload", "")[:60]))
        for k, v in d["kernels"].items(): print("   %-16s %7.4f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["GBps"], v["frac"]))
