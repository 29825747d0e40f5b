"""Times the tcgen05 attention core at the bench shape (252 hypotheses) and reports TFLOP/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from foundationpose_cpp_b200 import capi

n = int(sys.argv[1]) if len(sys.argv) > 1 else 252
qkv = torch.randn(n, 400, 1536, device="cuda").half()
out = torch.empty(n, 400, 512, device="cuda", dtype=torch.float16)
ms = capi.debug_attention(qkv.data_ptr(), out.data_ptr(), n, 20)
fl = 2.0 * n * 4 * 2 * 400 * 400 * 128
print(f"attention n={n}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s (algorithmic)")
