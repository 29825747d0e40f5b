"""Parity of the tcgen05 implicit-GEMM kernel (csrc/conv_gemm.cu) against plain PyTorch fp32 on
the same fp16-rounded operands.  Tolerance: the kernel accumulates in fp32 and rounds the result
to fp16 once, so |err| <= 2^-10 * |ref| + a small absolute term for cancellation.

Every case runs for every tile width (block_n 64 / 128 / 256, where it divides cout) x {single CTAs,
CTA pairs issuing tcgen05.mma.cta_group::2}: that reaches each of the ten conv_gemm_kernel<BN, PAIR, PER_TAP>
instantiations, in particular <256, pair, per-tap> and <128, pair, per-group>, the ones a 252-hypothesis
Register runs for 80 % of its time (a 5-hypothesis forward would pick 64-wide unpaired tiles on its own)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from foundationpose_cpp_b200 import capi, layouts

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[1, 3], ids=["cta", "pair_umma2"])
def cluster(request):
    return request.param


@pytest.fixture(params=[0, 64, 128, 256], ids=["bn_auto", "bn64", "bn128", "bn256"])
def block_n(request):
    return request.param


def _bn_ok(block_n, cout):
    if block_n and cout % block_n:
        pytest.skip(f"block_n {block_n} does not divide cout {cout}")


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _check(out, ref, what):
    from conftest import record_err
    out = out.float()
    err = (out - ref).abs()
    tol = 2e-3 * ref.abs() + 2e-2
    bad = err > tol
    assert not bad.any(), (f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err "
                           f"{float(err.max()):.4g}, ref absmax {float(ref.abs().max()):.4g}")
    record_err(f"conv_gemm/{what}", float(err.max()), float(tol.max()), f"ref absmax {float(ref.abs().max()):.3g}")


def _rand(*shape, scale=1.0):
    return (torch.randn(*shape, device="cuda") * scale).half()


def _desc(**kw):
    d = capi.ConvDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    return d


@pytest.mark.parametrize("rows,k,cout,relu", [(1000, 512, 64, 1), (1000, 512, 128, 0), (777, 512, 256, 1),
                                              (4000, 512, 1536, 0), (128, 64, 512, 0), (50000, 512, 512, 1)])
def test_linear(rows, k, cout, relu, cluster, block_n):
    _bn_ok(block_n, cout)
    torch.manual_seed(0)
    x = _rand(rows, k)
    w = _rand(cout, k, scale=k ** -0.5)
    b = torch.randn(cout, device="cuda")
    out = torch.full((rows, cout), float("nan"), device="cuda", dtype=torch.float16)
    d = _desc(in_=_ptr(x), n_img=1, planes=1, in_h=1, in_w=rows, cin=k, out_h=1, out_w=rows, bw=128, bh=1, bn=1,
              tap_kind=capi.TAPS_1X1, weight=_ptr(w), cout=cout, bias=_ptr(b), out=_ptr(out), relu=relu, cluster=cluster,
              block_n=block_n)
    capi.debug_conv_gemm(d)
    ref = x.float() @ w.float().t() + b
    if relu:
        ref = ref.relu()
    _check(out, ref, f"linear {rows}x{k}x{cout} bn{block_n} cl{cluster}")


@pytest.mark.parametrize("n,hw,cin,cout,tile,residual,post", [
    (5, 40, 128, 128, (8, 8, 2), True, False),
    (3, 40, 256, 256, (8, 8, 2), True, False),
    (3, 20, 512, 512, (4, 4, 8), True, True),
    (2, 80, 64, 64, (16, 8, 1), False, False),
    # the tile shapes picked for 20 x 20 maps with 1 .. 4 images (Track, N = 1 / 2): partial tiles in x and y
    (1, 20, 512, 512, (16, 8, 1), True, True),
    (2, 20, 512, 512, (8, 8, 2), True, False),
    (3, 20, 256, 512, (8, 8, 2), False, False),
])
def test_conv3x3_s1(n, hw, cin, cout, tile, residual, post, cluster, block_n):
    _bn_ok(block_n, cout)
    torch.manual_seed(1)
    x = _rand(n, hw, hw, cin)
    w = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5)
    b = torch.randn(cout, device="cuda")
    res = _rand(n, hw, hw, cout) if residual else None
    pe = torch.randn(hw * hw, cout, device="cuda") if post else None
    out = torch.full((n, hw, hw, cout), float("nan"), device="cuda", dtype=torch.float16)
    wp = layouts.pack_conv3x3(w)
    d = _desc(in_=_ptr(x), n_img=n, planes=1, in_h=hw, in_w=hw, cin=cin, out_h=hw, out_w=hw, bw=tile[0], bh=tile[1],
              bn=tile[2], tap_kind=capi.TAPS_3X3_S1, weight=_ptr(wp), cout=cout, bias=_ptr(b), residual=_ptr(res),
              post_add=_ptr(pe), out=_ptr(out), relu=1, cluster=cluster, block_n=block_n)
    capi.debug_conv_gemm(d)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1)
    if residual:
        ref = ref + res.float()
    ref = ref.relu()
    if post:
        ref = ref + pe.view(1, hw, hw, cout)
    _check(out, ref, f"conv3x3 s1 {hw}x{hw} {cin}->{cout} bn{block_n} cl{cluster}")


@pytest.mark.parametrize("n,hw,cin,cout,tile", [(3, 80, 64, 128, (8, 8, 2)), (5, 40, 256, 512, (4, 4, 8)),
                                                (1, 40, 256, 512, (16, 8, 1)), (2, 40, 256, 512, (8, 8, 2))])
def test_conv3x3_s2_phase_in(n, hw, cin, cout, tile, cluster, block_n):
    _bn_ok(block_n, cout)
    torch.manual_seed(2)
    x = _rand(n, hw, hw, cin)
    xs = layouts.phase_split(x)
    w = _rand(cout, cin, 3, 3, scale=(9 * cin) ** -0.5)
    b = torch.randn(cout, device="cuda")
    ho = hw // 2
    out = torch.full((n, ho, ho, cout), float("nan"), device="cuda", dtype=torch.float16)
    d = _desc(in_=_ptr(xs), n_img=n, planes=4, in_h=ho, in_w=ho, cin=cin, out_h=ho, out_w=ho, bw=tile[0], bh=tile[1],
              bn=tile[2], tap_kind=capi.TAPS_3X3_S2_PHASE, weight=_ptr(layouts.pack_conv3x3(w)), cout=cout,
              bias=_ptr(b), out=_ptr(out), relu=1, cluster=cluster, block_n=block_n)
    capi.debug_conv_gemm(d)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=1).permute(0, 2, 3, 1).relu()
    _check(out, ref, f"conv3x3 s2 {hw}x{hw} {cin}->{cout} bn{block_n} cl{cluster}")


def test_stem7_and_phase_split_out(cluster):
    # cout = 64: only the 64-wide tile exists for the stem
    torch.manual_seed(3)
    n = 3
    x = _rand(n, 160, 160, 6)
    buf = layouts.pad_stem_input(x)
    w = _rand(64, 6, 7, 7, scale=(49 * 6) ** -0.5)
    b = torch.randn(64, device="cuda")
    out = torch.full((n, 4, 40, 40, 64), float("nan"), device="cuda", dtype=torch.float16)
    pitch_b = layouts.STEM_PITCH * layouts.PACK_C * 2
    d = _desc(in_=_ptr(buf), n_img=n, planes=2, in_h=80, in_w=80, cin=64, stride_w=32, stride_plane=pitch_b,
              stride_h=2 * pitch_b, stride_img=160 * pitch_b, out_h=80, out_w=80, bw=16, bh=8, bn=1,
              tap_kind=capi.TAPS_STEM7_ROWS, weight=_ptr(layouts.pack_stem7(w)), cout=64, bias=_ptr(b),
              out=_ptr(out), relu=1, out_phase_split=1, cluster=cluster)
    capi.debug_conv_gemm(d)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=3).permute(0, 2, 3, 1).relu()
    _check(layouts.phase_merge(out), ref, f"stem 7x7 s2 cl{cluster}")


@pytest.mark.parametrize("n", [1, 3, 37])
def test_stem_direct_convolution(n):
    """csrc/stem_conv.cu: the stem as a direct tcgen05 convolution whose A operand is read in place from the
    parity-split input patch (no-swizzle K-major descriptors, two filter taps per MMA).  Against F.conv2d fp32 on the
    same fp16 operands; every output of every tile position (image borders = TMA out-of-bounds zero fill), 1 image
    (50 tiles, fewer CTAs than SMs), 3 and 37 images (1 850 tiles on 296 persistent CTAs: ring wrap-around)."""
    torch.manual_seed(30 + n)
    x = _rand(n, 160, 160, 6)
    buf = layouts.pad_stem_input(x)
    w = _rand(64, 6, 7, 7, scale=(49 * 6) ** -0.5)
    b = torch.randn(64, device="cuda")
    out = torch.full((n, 4, 40, 40, 64), float("nan"), device="cuda", dtype=torch.float16)
    wf = layouts.pack_stem7(w).contiguous()
    capi.debug_stem_conv(buf.data_ptr(), n, wf.data_ptr(), b.data_ptr(), out.data_ptr())
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, stride=2, padding=3).permute(0, 2, 3, 1).relu()
    _check(layouts.phase_merge(out), ref, f"stem direct conv n{n}")


def test_concat_halves(cluster, block_n):
    """Images n >= img_mod land in the same pixel at a channel offset (a||b feature concat)."""
    _bn_ok(block_n, 128)
    torch.manual_seed(4)
    n, hw, c = 6, 40, 128
    x = _rand(n, hw, hw, c)
    w = _rand(c, c, 3, 3, scale=(9 * c) ** -0.5)
    b = torch.randn(c, device="cuda")
    out = torch.full((n // 2, hw, hw, 2 * c), float("nan"), device="cuda", dtype=torch.float16)
    d = _desc(in_=_ptr(x), n_img=n, planes=1, in_h=hw, in_w=hw, cin=c, out_h=hw, out_w=hw, bw=8, bh=8, bn=2,
              tap_kind=capi.TAPS_3X3_S1, weight=_ptr(layouts.pack_conv3x3(w)), cout=c, bias=_ptr(b), out=_ptr(out),
              relu=1, out_ld=2 * c, img_mod=n // 2, half_coff=c, cluster=cluster, block_n=block_n)
    capi.debug_conv_gemm(d)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1).permute(0, 2, 3, 1).relu()
    ref = torch.cat([ref[: n // 2], ref[n // 2:]], dim=-1)
    _check(out, ref, f"concat bn{block_n} cl{cluster}")
