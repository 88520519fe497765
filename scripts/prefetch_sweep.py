"""L2 prefetch distance sweep for the fused kernel (run on a B200):  python scripts/prefetch_sweep.py"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for pf in [int(x) for x in (sys.argv[1:] or "0 2816 5632 8448".split())]:
    env = dict(os.environ, VSB_PREFETCH_KB=str(pf))
    print("== prefetch KB", pf, flush=True)
    subprocess.run([sys.executable, os.path.join(root, "scripts", "vec_sweep.py"), "quick"], env=env)
