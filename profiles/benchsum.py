"""Print the per-kernel table of bench.py JSON lines.
usage: python profiles/benchsum.py a.json [b.json ...]   (or JSON lines on stdin when no file is given)"""
import json, sys
texts = [open(f).read() for f in sys.argv[1:]] if len(sys.argv) > 1 else [sys.stdin.read()]
for text in texts:
    for line in text.strip().splitlines():
        if not line.startswith('{'): continue
        d = json.loads(line)
        print("ms/step %.3f  img/s %.1f  e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
        for k, v in d["kernels"].items(): print("   %-16s %7.4f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["GBps"], v["frac"]))
