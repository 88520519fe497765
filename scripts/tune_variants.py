"""Build (here) and time (on the GPU box) register-budget variants of the D3Q19 fused kernel.

    python scripts/tune_variants.py build          # nvcc, no GPU needed -> build_variants/*.so
    python scripts/tune_variants.py run            # on a B200: scripts/vec_sweep.py once per variant
"""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, root)
VARIANTS = {"t256_c3": ("VSB_STEP3D_THREADS=256", "VSB_STEP3D_CTAS=3"),
            "t128_c5": ("VSB_STEP3D_THREADS=128", "VSB_STEP3D_CTAS=5"),
            "t128_c4": ("VSB_STEP3D_THREADS=128", "VSB_STEP3D_CTAS=4")}
vdir = os.path.join(root, "build_variants")

if sys.argv[1] == "build":
    from vivsim_b200 import _build
    os.makedirs(vdir, exist_ok=True)
    for name, defs in VARIANTS.items():
        print(_build.build(out=os.path.join(vdir, name + ".so"), defines=defs))
else:
    for name in ["product"] + list(VARIANTS):
        env = dict(os.environ)
        if name != "product": env["VIVSIM_B200_LIB"] = os.path.join(vdir, name + ".so")
        print("==", name, flush=True)
        subprocess.run([sys.executable, os.path.join(root, "scripts", "vec_sweep.py"), "3d"], env=env)
