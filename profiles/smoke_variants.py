"""Does a library build survive the batched kernels? (run on the GPU box) python profiles/smoke_variants.py lib1.so ..."""
import os, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
if os.environ.get("NDT_SMOKE_CHILD"):
    import numpy as np
    import bench
    from ndt_slam_b200 import capi
    wl = bench.build_c4(65536)
    g = capi.Ndt(capi.default_params(resolution=0.5))
    g.set_target(wl["tgt"]); g.set_source(wl["src"])
    try:
        r = g.align_batch(np.ascontiguousarray(wl["hyp"][:4096])); print("c4 ok", int(r["evals"].sum()), flush=True)
    except Exception as e:
        print("c4 FAIL", str(e)[:120], flush=True); sys.exit(0)
    c5 = bench.build_c5(0, 256)
    for sched in (capi.PAIRS_WARP, capi.PAIRS_CTA):
        g5 = capi.Ndt(capi.default_params(resolution=0.5, pairs_schedule=sched))
        try:
            r5 = g5.match_pairs(c5["src"], c5["so"], c5["tgt"], c5["to"], np.zeros((256, 3)), 256, source_leaf=0.05); print("pairs ok", sched, int(r5["evals"].sum()), flush=True)
        except Exception as e:
            print("pairs FAIL", sched, str(e)[:120], flush=True); sys.exit(0)
else:
    for lib in sys.argv[1:]:
        env = dict(os.environ, NDT_SMOKE_CHILD="1", NDT_B200_LIB=str((ROOT / lib).resolve()))
        cp = subprocess.run([sys.executable, __file__], env=env, capture_output=True, text=True, timeout=600)
        print(lib, "|", " ; ".join(l for l in cp.stdout.splitlines()), "| rc", cp.returncode, flush=True)
