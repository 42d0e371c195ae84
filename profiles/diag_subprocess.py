import subprocess, sys, json
for mode in ("capture", "file"):
    if mode == "capture":
        cp = subprocess.run([sys.executable, "bench.py", "--extras-only", "--c2-scans", "300"], capture_output=True, text=True)
        out = cp.stdout
    else:
        with open("/tmp/x.json", "w") as f:
            subprocess.run([sys.executable, "bench.py", "--extras-only", "--c2-scans", "300"], stdout=f)
        out = open("/tmp/x.json").read()
    d = json.loads(out.strip().splitlines()[-1])
    print(mode, d["C2"]["ms_per_scan"], d["C2"]["stage_ms_per_scan"]["device_grid_kernels"], d["C2"]["stage_ms_per_scan"]["set_target_call"])
