import json,sys
for line in sys.stdin.read().strip().splitlines():
    if not line.startswith('{'): continue
    d=json.loads(line)
    print("ms/step %.3f  img/s %.1f  e2e %.1f" % (d["ms_per_step"], d["value"], d["e2e"]["value"]))
    for k,v in d["kernels"].items(): print("   %-16s %7.4f ms  %6.0f GB/s  frac %.3f" % (k, v["ms"], v["GBps"], v["frac"]))
