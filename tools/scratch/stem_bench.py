import os, sys, json
sys.path.insert(0, "/root/repo")
import torch
from foundationpose_cpp_b200 import capi, layouts
for n in (64, 126, 504):
    x = (torch.randn(n, 160, 160, 6, device="cuda")).half()
    buf = layouts.pad_stem_input(x)
    w = (torch.randn(64, 6, 7, 7, device="cuda") * 0.06).half()
    b = torch.randn(64, device="cuda")
    out = torch.empty((n, 4, 40, 40, 64), device="cuda", dtype=torch.float16)
    wf = layouts.pack_stem7(w).contiguous()
    ms = capi.debug_stem_conv(buf.data_ptr(), n, wf.data_ptr(), b.data_ptr(), out.data_ptr(), 20)
    fl = 2.0 * n * 6400 * 64 * 294
    print(json.dumps({"imgs": n, "ms": round(ms, 4), "tflops_algorithmic": round(fl / ms / 1e9, 1)}))
