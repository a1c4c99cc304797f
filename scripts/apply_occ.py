"""AR-FPN apply backward under ARFE_APPLY_OCC (resident CTAs per SM the kernel is compiled for)."""
import os, sys, subprocess
for occ in ("4", "5", "6"):
    env = dict(os.environ, ARFE_APPLY_OCC=occ)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "40", "--warmup", "8"], env=env, capture_output=True, text=True).stdout
    import json
    d = json.loads(out.strip().splitlines()[-1])
    print(occ, {k: v["ms"] for k, v in d["kernels"].items() if "apply" in k}, d["ms_per_step"])
