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
if os.environ.get("FPB200_ATTN_PHASES") and os.environ.get("FPB200_ATTN_V2", "0") == "1":
    import ctypes as C
    buf = (C.c_ulonglong * 22)()
    capi.check(capi.lib().fpb200_debug_attention_phases(buf))
    items = max(1, buf[7])
    mn = ["wait K/Q", "-", "-", "wait V", "wait P", "wait O free", "issue/other"]
    print(f"attention2, CTA 0, {items} items; SM clocks per item")
    print("  MMA thread B: " + ", ".join(f"{nm} {buf[i] / items:.0f}" for i, nm in enumerate(mn) if nm != "-") + f"; total {sum(buf[0:7]) / items:.0f}")
    print("  MMA thread A: " + ", ".join(f"{nm} {buf[18 + i] / items:.0f}" for i, nm in enumerate(["wait K/Q", "wait S_A free", "issue/other"])))
    print("  set B (exp): " + ", ".join(f"{nm} {buf[8 + i] / items:.0f}" for i, nm in enumerate(["waits (max, S_B)", "tcgen05.ld + wait", "exp2 / convert / sum", "tcgen05.st + wait, fence, arrive"])))
    an = ["wait S_A", "pass A compute / other", "wait Q free", "wait O + sum", "epilogue", "(pre-epilogue)"]
    print("  set A: " + ", ".join(f"{nm} {buf[12 + i] / items:.0f}" for i, nm in enumerate(an)))
elif os.environ.get("FPB200_ATTN_PHASES"):
    import ctypes as C
    buf = (C.c_ulonglong * 22)()
    capi.check(capi.lib().fpb200_debug_attention_phases(buf))
    names = ["wait S", "pass 1 (max)", "pass 2 (exp)", "wait O", "epi tmem->smem", "store issue"]
    for kind, o in (("full tiles", 0), ("tail tiles", 7)):
        items = max(1, buf[o + 6])
        print(f"CTA 0, {kind} (SM clocks per item): " + ", ".join(f"{nm} {buf[o + i] / items:.0f}" for i, nm in enumerate(names))
              + f"; items {buf[o + 6]}, total/item {sum(buf[o:o + 6]) / items:.0f}")
    n_it = max(1, buf[6] + buf[13])
    mn = ["wait Q", "wait K", "wait O drained", "wait V", "wait P steps", "issue/other"]
    print("CTA 0, MMA thread (SM clocks per item): " + ", ".join(f"{nm} {buf[16 + i] / n_it:.0f}" for i, nm in enumerate(mn)))
