"""Phase-skip decomposition of the fused NonLocal2D attention (profile build only):
ARFE_NL_DBG bits: 1 no K / V copies, 2 no softmax arithmetic, 4 no MMAs, 16 no row-max exchange between the
column halves, 32 no ex2, 64 no P store, 128 no S load, 1024 two key steps only (the fixed cost of a CTA).
Results are garbage.  The call is timed as a CUDA graph replay (the Python wrapper costs ~40 us of host time).
usage: python scripts/nonlocal_knobs.py [bits ...]"""
import os
import subprocess
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
SNIP = r'''
import sys, torch
sys.path.insert(0, %r)
import arfe_b200 as A
B, D, H, W = 2, 256, 50, 84
ts = [torch.randn(B, D, H, W, device="cuda").contiguous(memory_format=torch.channels_last) * (0.25 if i < 2 else 1) for i in range(3)]
for ns in (1, 2):
    for _ in range(3): A.nonlocal_attention(*ts, 1.0, ns)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph(); st = torch.cuda.Stream(); st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            A.nonlocal_attention(*ts, 1.0, ns)
    torch.cuda.current_stream().wait_stream(st)
    gr.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): gr.replay()
    e1.record(); torch.cuda.synchronize()
    print("nsplit", ns, "graph %%.1f us" %% (e0.elapsed_time(e1) / 20 * 1e3))
'''
for dbg in [int(a) for a in sys.argv[1:]] or (0, 1, 4, 5):
    env = dict(os.environ, ARFE_B200_LIB=os.path.join(ROOT, "arfe_b200", "libarfe_b200_prof.so"), ARFE_NL_DBG=str(dbg))
    r = subprocess.run([sys.executable, "-c", SNIP % ROOT], env=env, capture_output=True, text=True, timeout=300)
    print("ARFE_NL_DBG", dbg, "|", " | ".join(r.stdout.strip().splitlines()), r.stderr.strip()[-300:] if r.returncode else "", flush=True)
