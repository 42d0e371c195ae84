import sys, runpy
sys.path.insert(0, "/root/repo")
import torch, numpy as np
import bench
from ndt_slam_b200 import capi
# mimic the bench process before the extras: C4 batch on a torch stream + flush buffer + pinned buffers
wl = bench.build_c4(65536)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
prm = capi.default_params(resolution=0.5, device=0, stream=stream.cuda_stream)
g = capi.Ndt(prm); g.set_target(wl["tgt"]); g.set_source(wl["src"])
hyp = np.ascontiguousarray(wl["hyp"][:65536]); d_hyp = torch.from_numpy(hyp).cuda()
d_res = torch.zeros(65536 * capi.RESULT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for _ in range(3):
    flush.zero_(); g.align_batch(d_hyp.data_ptr(), n=65536, space=capi.MEM_DEVICE, out=d_res.data_ptr())
torch.cuda.synchronize()
print("in-process after C4:")
sys.argv = ["c2_series.py", "300"]
runpy.run_path("/root/repo/profiles/c2_series.py", run_name="__main__")
