"""Micro-benchmark of the tcgen05 implicit-GEMM kernel on the network's layer shapes at N hypotheses.
Prints one line per layer: ms, TFLOP/s, fraction of the measured bf16 peak."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foundationpose_cpp_b200 import capi, layouts  # noqa: E402


def ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 252
    rep = 5
    peak = 1643.8
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:
        pass
    cases = [
        # name, imgs, hw_in, cin, cout, tile, kind
        ("stem7 6->64 s2", 2 * n, 160, 64, 64, (16, 8, 1), capi.TAPS_STEM7_ROWS),
        ("c3 64->128 s2", 2 * n, 80, 64, 128, (8, 8, 2), capi.TAPS_3X3_S2_PHASE),
        ("c3 128->128", 2 * n, 40, 128, 128, (8, 8, 2), capi.TAPS_3X3_S1),
        ("c3 256->256", n, 40, 256, 256, (8, 8, 2), capi.TAPS_3X3_S1),
        ("c3 256->512 s2", n, 40, 256, 512, (4, 4, 8), capi.TAPS_3X3_S2_PHASE),
        ("c3 512->512", n, 20, 512, 512, (4, 4, 8), capi.TAPS_3X3_S1),
        ("lin 512->1536", 1, n * 400, 512, 1536, (128, 1, 1), capi.TAPS_1X1),
        ("lin 512->512", 1, n * 400, 512, 512, (128, 1, 1), capi.TAPS_1X1),
        ("lin 512->512 +res", 1, n * 400, 512, 512, (128, 1, 1), capi.TAPS_1X1),
    ]
    only = os.environ.get("FPB_CASE")
    for ci, (name, imgs, hw, cin, cout, tile, kind) in enumerate(cases):
        if only is not None and str(ci) not in only.split(","):
            continue
        d = capi.ConvDesc()
        if kind == capi.TAPS_STEM7_ROWS:
            x = torch.randn(imgs * 160 * layouts.STEM_PITCH * 8 + 4096, device="cuda").half()
            pitch_b = layouts.STEM_PITCH * 16
            d.planes, d.in_h, d.in_w, d.cin = 2, 80, 80, 64
            d.stride_w, d.stride_plane, d.stride_h, d.stride_img = 32, pitch_b, 2 * pitch_b, 160 * pitch_b
            ho = wo = 80
            taps = 7
        elif kind == capi.TAPS_3X3_S2_PHASE:
            ho = wo = hw // 2
            x = torch.randn(imgs, 4, ho, wo, cin, device="cuda").half()
            d.planes, d.in_h, d.in_w, d.cin = 4, ho, wo, cin
            taps = 9
        elif kind == capi.TAPS_3X3_S1:
            ho = wo = hw
            x = torch.randn(imgs, hw, hw, cin, device="cuda").half()
            d.planes, d.in_h, d.in_w, d.cin = 1, hw, hw, cin
            taps = 9
        else:
            ho, wo = 1, hw
            x = torch.randn(hw, cin, device="cuda").half()
            d.planes, d.in_h, d.in_w, d.cin = 1, 1, hw, cin
            taps = 1
        w = (torch.randn(cout, taps * d.cin, device="cuda") * 0.02).half()
        b = torch.zeros(cout, device="cuda")
        out = torch.empty(imgs * ho * wo * cout, device="cuda", dtype=torch.float16)
        d.in_, d.n_img, d.out_h, d.out_w = ptr(x), imgs, ho, wo
        d.bw, d.bh, d.bn, d.tap_kind = tile[0], tile[1], tile[2], kind
        d.weight, d.cout, d.bias, d.out, d.relu = ptr(w), cout, ptr(b), ptr(out), 1
        if name.endswith("+res"):  # out-proj / FFN2 of the encoder layers: residual sum in the epilogue, no ReLU
            res = torch.randn(hw, cout, device="cuda").half()
            d.residual, d.res_ld, d.relu = ptr(res), cout, 0
        ms = capi.debug_conv_gemm(d, rep)
        flops = 2.0 * imgs * ho * wo * cout * taps * d.cin
        tf = flops / ms / 1e9
        print(f"{name:18s} M={imgs*ho*wo:8d} N={cout:4d} K={taps*d.cin:5d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  "
              f"{tf/peak*100:5.1f}% of measured burst peak", flush=True)


if __name__ == "__main__":
    main()
