"""A/B timing of library builds on the headline workloads (run on the GPU box):
    python profiles/ab_c4.py [--pairs] [--c1] lib1.so lib2.so ...
Each library runs in its own process (NDT_B200_LIB). Prints per library: C4 ms/step (65,536 hypotheses, median of 5 after
3 warm-ups, L2 flushed between steps), G point-evals/s, and a digest of the results (must be equal across builds that
claim bit-identical arithmetic); with --pairs the C5 figures (8,192 and 1,024 pairs), with --c1 the C1 single-match latency."""
import hashlib
import json
import os
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(args):
    import numpy as np
    import torch

    import bench
    from ndt_slam_b200 import capi

    out = {"lib": os.environ.get("NDT_B200_LIB", "default")}
    wl = bench.build_c4(65536)
    stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
    prm = capi.default_params(resolution=0.5, stream=stream.cuda_stream)
    g = capi.Ndt(prm)
    g.set_target(wl["tgt"]); g.set_source(wl["src"])
    n = wl["hyp"].shape[0]
    d_h = torch.from_numpy(wl["hyp"]).cuda()
    d_r = torch.zeros(n * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ms = []
    for it in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(stream); g.align_batch(d_h.data_ptr(), n=n, space=capi.MEM_DEVICE, out=d_r.data_ptr()); b.record(stream)
        torch.cuda.synchronize()
        if it >= 3:
            ms.append(a.elapsed_time(b))
    res = np.frombuffer(d_r.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
    pe = int(res["point_evals"].sum())
    if os.environ.get("NDT_AB_DUMP"):
        (ROOT / "gpurun_out" / "r2").mkdir(parents=True, exist_ok=True)
        np.save(ROOT / "gpurun_out" / "r2" / f"c4_{Path(out['lib']).stem}.npy", res)
    out["c4_passes_run_frac"] = float(res["passes_run"].sum() / max(res["evals"].sum(), 1))
    out["c4_ms"] = float(np.median(ms)); out["c4_gpe"] = pe / (np.median(ms) * 1e-3) / 1e9
    out["c4_digest"] = hashlib.sha256(res["pose"].tobytes() + res["score"].tobytes() + res["iters"].tobytes() + res["evals"].tobytes()).hexdigest()[:12]
    d_o = torch.zeros((n, 14), dtype=torch.float64, device="cuda")
    sw = []
    for it in range(4):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record(stream); g.eval_batch(d_h.data_ptr(), n=n, want_hessian=False, space=capi.MEM_DEVICE, out=d_o.data_ptr()); b.record(stream)
        torch.cuda.synchronize(); sw.append(a.elapsed_time(b))
    out["sweep_ms"] = float(min(sw))
    if "--team" in args:
        # warps per match (ndt_params.align_team) x hypotheses per call: ms per call, median of 5 after 2 warm-ups, L2 flushed
        table = {}
        for nh in (65536, 32768, 16384, 8192, 4096, 2048, 512):
            for team in (1, 2, 4, 8):
                gt = capi.Ndt(capi.default_params(resolution=0.5, stream=stream.cuda_stream, align_team=team))
                gt.set_target(wl["tgt"]); gt.set_source(wl["src"])
                tt = []
                for it in range(7):
                    flush.zero_()
                    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                    a.record(stream); gt.align_batch(d_h.data_ptr(), n=nh, space=capi.MEM_DEVICE, out=d_r.data_ptr(), want_fitness=False); b.record(stream)
                    torch.cuda.synchronize()
                    if it >= 2:
                        tt.append(a.elapsed_time(b))
                table[f"{nh}x{team}"] = round(float(np.median(tt)), 4)
        out["team_ms"] = table
    if "--pairs" in args:
        for npairs, sched in ((8192, capi.PAIRS_WARP), (8192, capi.PAIRS_CTA), (2048, capi.PAIRS_CTA), (1024, capi.PAIRS_CTA)):
            c5 = bench.build_c5(0, npairs)
            prm5 = capi.default_params(resolution=0.5, stream=stream.cuda_stream, pairs_schedule=sched)
            g5 = capi.Ndt(prm5)
            tag = f"{npairs}" + {capi.PAIRS_WARP: "warp", capi.PAIRS_CTA: "cta", capi.PAIRS_AUTO: ""}[sched]
            d_s, d_t = torch.from_numpy(c5["src"]).cuda(), torch.from_numpy(c5["tgt"]).cuda()
            d_g = torch.zeros((npairs, 3), dtype=torch.float64, device="cuda")
            d_r5 = torch.zeros(npairs * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
            t5 = []
            for it in range(6):
                flush.zero_()
                a, b = torch.cuda.Event(True), torch.cuda.Event(True)
                a.record(stream)
                g5.match_pairs(d_s.data_ptr(), c5["so"], d_t.data_ptr(), c5["to"], d_g.data_ptr(), npairs, source_leaf=0.05, space=capi.MEM_DEVICE, out=d_r5.data_ptr())
                b.record(stream); torch.cuda.synchronize()
                if it >= 2:
                    t5.append(a.elapsed_time(b))
            r5 = np.frombuffer(d_r5.cpu().numpy().tobytes(), dtype=capi.RESULT_DTYPE)
            if os.environ.get("NDT_AB_DUMP"):
                np.save(ROOT / "gpurun_out" / "r2" / f"c5_{tag}_{Path(out['lib']).stem}.npy", r5)
            out[f"c5_{tag}_ms"] = float(np.median(t5))
            out[f"c5_{tag}_digest"] = hashlib.sha256(r5["pose"].tobytes() + r5["fitness"].tobytes() + r5["evals"].tobytes()).hexdigest()[:12]
    if "--c1" in args:
        sys.path.insert(0, str(ROOT / "tests"))
        import ndt_common as common
        pb = common.c1_problem()
        g1 = capi.Ndt(capi.default_params(resolution=0.5))
        ks, bs = [], []
        for it in range(30):
            g1.set_target(pb["tgt"]); bs.append(g1.last_kernel_ms()); g1.set_source(pb["src"])
            r = g1.align(pb["guess"]); ks.append(g1.last_kernel_ms())
        out["c1_match_ms"] = float(np.median(ks[5:])); out["c1_grid_ms"] = float(np.median(bs[5:]))
        out["c1_digest"] = hashlib.sha256(bytes(r)).hexdigest()[:12]
    print("AB " + json.dumps(out), flush=True)


if __name__ == "__main__":
    if os.environ.get("NDT_AB_CHILD"):
        child(sys.argv[1:])
    else:
        flags = [a for a in sys.argv[1:] if a.startswith("--")]
        libs = [a for a in sys.argv[1:] if not a.startswith("--")] or [""]
        for lib in libs:
            env = dict(os.environ, NDT_AB_CHILD="1")
            if lib:
                env["NDT_B200_LIB"] = str((ROOT / lib).resolve()) if not os.path.isabs(lib) else lib
            cp = subprocess.run([sys.executable, __file__, *flags], env=env, capture_output=True, text=True, timeout=900)
            lines = [l for l in cp.stdout.splitlines() if l.startswith("AB ")]
            print(lines[-1] if lines else f"AB-FAIL {lib}: rc={cp.returncode} {cp.stderr[-600:]}", flush=True)
