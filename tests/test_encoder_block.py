"""Parity of the fused encoder-layer tail (csrc/encoder_block.cu: out-projection + residual + LayerNorm1 + FFN +
residual + LayerNorm2 + token mean, one tcgen05 kernel) against plain PyTorch fp32 on the same fp16 operands, stage by
stage (the kernel can dump LayerNorm1 output, FFN activation and LayerNorm2 output) and for the token mean.

The kernel rounds the two pre-LayerNorm sums and h to fp16 (as the unfused layers stored them); the reference does
the same so that what is compared is arithmetic, not storage.  Tolerances: stages 2e-2 absolute on O(1) LayerNorm
outputs (one fp16 ulp of an input can move a normalised value by ~1e-3; GEMM K = 512), token mean 3e-3."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

from foundationpose_cpp_b200 import capi

pytestmark = pytest.mark.gpu


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _reference(att, x0, w, b, ln):
    v1 = (att.float() @ w[0].float().t() + b[0] + x0.float()).half().float()
    x1 = F.layer_norm(v1, (512,), ln[0], ln[1], 1e-5).half().float()
    h = (x1 @ w[1].float().t() + b[1]).relu().half().float()
    v2 = (h @ w[2].float().t() + b[2] + x1).half().float()
    y = F.layer_norm(v2, (512,), ln[2], ln[3], 1e-5)
    return x1, h, y


@pytest.mark.parametrize("mode", [1, 2], ids=["cta", "pair_umma2"])
@pytest.mark.parametrize("n_hyp", [1, 3, 37])
def test_encoder_block_stages_and_mean(n_hyp, mode, record):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(7 + n_hyp)
    rows = n_hyp * 400
    att = (torch.randn(rows, 512, device="cuda") * 0.7).half()
    x0 = (torch.randn(rows, 512, device="cuda") * 1.3 + 0.2).half()
    w = [(torch.randn(512, 512, device="cuda") * 512 ** -0.5).half() for _ in range(3)]
    b = [torch.randn(512, device="cuda") * 0.1 for _ in range(3)]
    ln = [1 + 0.1 * torch.randn(512, device="cuda"), 0.05 * torch.randn(512, device="cuda"),
          1 + 0.1 * torch.randn(512, device="cuda"), 0.05 * torch.randn(512, device="cuda")]
    partial = torch.full((n_hyp * 25, 512), float("nan"), device="cuda")
    mean = torch.full((n_hyp, 512), float("nan"), device="cuda")
    dbg = [torch.full((rows, 512), float("nan"), device="cuda", dtype=torch.float16) for _ in range(3)]
    d = capi.EncoderDesc()
    d.att, d.x0 = _ptr(att), _ptr(x0)
    d.w_out, d.w1, d.w2 = _ptr(w[0]), _ptr(w[1]), _ptr(w[2])
    d.b_out, d.b1, d.b2 = _ptr(b[0]), _ptr(b[1]), _ptr(b[2])
    d.ln1_g, d.ln1_b, d.ln2_g, d.ln2_b = (_ptr(t) for t in ln)
    d.partial, d.mean_out, d.n_hyp, d.mode = _ptr(partial), _ptr(mean), n_hyp, mode
    d.dbg_x1, d.dbg_h, d.dbg_y = (_ptr(t) for t in dbg)
    capi.debug_encoder_block(d)
    torch.cuda.synchronize()
    x1, h, y = _reference(att, x0, w, b, ln)
    tag = f"encoder_block n{n_hyp} mode{mode}"
    for name, got, ref, tol in (("LayerNorm1 out", dbg[0], x1, 2e-2), ("FFN activation", dbg[1], h, 2e-2),
                                ("LayerNorm2 out", dbg[2], y, 3e-2)):
        assert torch.isfinite(got.float()).all(), f"{name}: rows were not written"
        err = float((got.float() - ref).abs().max())
        record(f"{tag}/{name}", err, tol)
        assert err <= tol, f"{name}: max err {err:.4g}"
    ref_mean = y.view(n_hyp, 400, 512).mean(1)
    err = float((mean - ref_mean).abs().max())
    record(f"{tag}/token mean", err, 3e-3)
    assert err <= 3e-3, f"token mean: max err {err:.4g}"
    # the group sums are deterministic: a second launch reproduces them bit for bit
    mean2 = torch.empty_like(mean)
    d.mean_out = _ptr(mean2)
    capi.debug_encoder_block(d)
    torch.cuda.synchronize()
    assert torch.equal(mean, mean2)


def test_encoder_block_independent_of_batch_position():
    """A hypothesis's token mean must not depend on how many hypotheses precede it (sharded runs are bit-identical
    to unsharded ones): hypothesis 5 of a 9-hypothesis launch == the same rows launched alone."""
    torch.manual_seed(3)
    n = 9
    att = (torch.randn(n * 400, 512, device="cuda") * 0.7).half()
    x0 = (torch.randn(n * 400, 512, device="cuda") * 1.3).half()
    w = [(torch.randn(512, 512, device="cuda") * 512 ** -0.5).half() for _ in range(3)]
    b = [torch.randn(512, device="cuda") * 0.1 for _ in range(3)]
    ln = [1 + 0.1 * torch.randn(512, device="cuda"), 0.05 * torch.randn(512, device="cuda"),
          1 + 0.1 * torch.randn(512, device="cuda"), 0.05 * torch.randn(512, device="cuda")]

    def run(a, x, n_hyp, mode):
        partial = torch.zeros(n_hyp * 25, 512, device="cuda")
        mean = torch.zeros(n_hyp, 512, device="cuda")
        d = capi.EncoderDesc()
        d.att, d.x0 = _ptr(a), _ptr(x)
        d.w_out, d.w1, d.w2 = _ptr(w[0]), _ptr(w[1]), _ptr(w[2])
        d.b_out, d.b1, d.b2 = _ptr(b[0]), _ptr(b[1]), _ptr(b[2])
        d.ln1_g, d.ln1_b, d.ln2_g, d.ln2_b = (_ptr(t) for t in ln)
        d.partial, d.mean_out, d.n_hyp, d.mode = _ptr(partial), _ptr(mean), n_hyp, mode
        capi.debug_encoder_block(d)
        torch.cuda.synchronize()
        return mean

    full = run(att, x0, n, 2)
    lo, hi = 5 * 400, 6 * 400
    alone = run(att[lo:hi].contiguous(), x0[lo:hi].contiguous(), 1, 1)
    assert torch.equal(full[5:6], alone)
