"""Micro-benchmark of the fused encoder-layer tail (csrc/encoder_block.cu) at N hypotheses, both modes."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foundationpose_cpp_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 252
    repeat = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    rows = n * 400
    g = torch.Generator(device="cuda").manual_seed(0)
    att = (torch.randn(rows, 512, device="cuda", generator=g) * 0.5).half()
    x0 = torch.randn(rows, 512, device="cuda", generator=g).half()
    w = [(torch.randn(512, 512, device="cuda", generator=g) * 512 ** -0.5).half() for _ in range(3)]
    v = [torch.randn(512, device="cuda", generator=g) * 0.1 for _ in range(7)]
    partial, mean = torch.zeros(n * 25, 512, device="cuda"), torch.zeros(n, 512, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    out = {}
    for mode in (1, 2):
        d = capi.EncoderDesc()
        d.att, d.x0, d.w_out, d.w1, d.w2 = p(att), p(x0), p(w[0]), p(w[1]), p(w[2])
        d.b_out, d.b1, d.b2, d.ln1_g, d.ln1_b, d.ln2_g, d.ln2_b = (p(t) for t in v)
        d.partial, d.mean_out, d.n_hyp, d.mode = p(partial), p(mean), n, mode
        ms = capi.debug_encoder_block(d, repeat)
        fl = 3 * 2.0 * rows * 512 * 512
        clk = torch.zeros(24, dtype=torch.int64, device="cuda")
        d.dbg_clocks = p(clk)
        capi.debug_encoder_block(d, 0)
        torch.cuda.synchronize()
        c = clk.cpu().tolist()
        nb = max(c[6], 1)
        names_m = ["wait accempty", "wait att k-block", "wait x1 k-block", "wait h k-block", "wait weights", "issue"]
        names_e = ["wait slab", "wait acc G1", "wait acc G2", "wait acc G3", "passA G1", "passA G2", "passA G3",
                   "stats exchange", "LN1 pass", "LN2 pass"]
        out["cta" if mode == 1 else "pair"] = {
            "ms": ms, "tflops": fl / (ms * 1e-3) / 1e12, "blocks_cta0": c[6],
            "mma_clocks_per_block": {n: round(v / nb) for n, v in zip(names_m, c[:6])},
            "epilogue_clocks_per_block": {n: round(v / max(c[20], 1)) for n, v in zip(names_e, c[8:18])}}
    print(json.dumps({"n_hyp": n, **out}))


if __name__ == "__main__":
    main()
