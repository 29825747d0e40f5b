"""tcgen05 self-attention core (csrc/attention.cu) against a plain PyTorch fp32 reference of the same op:
softmax(Q K^T / sqrt(128)) V for 400 tokens, 4 heads x 128 dims per hypothesis (the op TensorRT runs
inside the refine / score engines of the reference, trt_core.cpp:367)."""
import os

import pytest
import torch

from foundationpose_cpp_b200 import capi

pytestmark = pytest.mark.gpu


def reference(qkv: torch.Tensor) -> torch.Tensor:
    n = qkv.shape[0]
    q, k, v = [t.float().view(n, 400, 4, 128).transpose(1, 2) for t in qkv.split(512, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * (128 ** -0.5)
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(n, 400, 512)


@pytest.fixture(params=["v2", "v1"])
def design(request, monkeypatch):
    """Both kernels: attention.cu (default) and attention2.cu (FPB200_ATTN_V2=1, read when a plan is built)."""
    monkeypatch.setenv("FPB200_ATTN_V2", "1" if request.param == "v2" else "0")
    return request.param


@pytest.mark.parametrize("n_hyp,scale", [(1, 1.0), (3, 1.0), (40, 0.5), (7, 3.0), (75, 1.0)])
def test_attention_matches_fp32_reference(n_hyp, scale, design):
    g = torch.Generator(device="cuda").manual_seed(n_hyp)
    qkv = (torch.randn(n_hyp, 400, 1536, device="cuda", generator=g) * scale).half()
    out = torch.full((n_hyp, 400, 512), float("nan"), device="cuda", dtype=torch.float16)
    capi.debug_attention(qkv.data_ptr(), out.data_ptr(), n_hyp)
    torch.cuda.synchronize()
    ref = reference(qkv)
    assert torch.isfinite(out).all()
    err = (out.float() - ref).abs().max().item()
    # fp16 P and fp16 output rounding: 2e-3 of the value range (|v| <~ 4 scale)
    assert err < 2e-3 * max(1.0, 4 * scale), err


def test_attention_structured_rows(design):
    """Each output row must depend on its own query only: one-hot-like keys make the answer known."""
    n = 2
    qkv = torch.zeros(n, 400, 1536, device="cuda", dtype=torch.float16)
    # key t of every head = 8 * e_(t mod 128); query t = 8 * e_((t + 1) mod 128); value t = t / 400 in all dims
    t = torch.arange(400, device="cuda")
    for h in range(4):
        qkv[:, t, 512 + h * 128 + (t % 128)] = 8.0
        qkv[:, t, h * 128 + ((t + 1) % 128)] = 8.0
    qkv[:, :, 1024:] = (t.float() / 400).half()[None, :, None]
    out = torch.empty(n, 400, 512, device="cuda", dtype=torch.float16)
    capi.debug_attention(qkv.data_ptr(), out.data_ptr(), n)
    torch.cuda.synchronize()
    ref = reference(qkv)
    assert (out.float() - ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("tpc", [1, 2, 4])
@pytest.mark.parametrize("n_hyp", [2, 40, 110])
def test_attention2_tiles_per_cta(n_hyp, tpc, monkeypatch):
    """attention2.cu: every split of a head's four query tiles over CTAs (1, 2 or 4 tiles per CTA), with one and with
    several (hypothesis, head) units per CTA (K / V hand-over between units)."""
    monkeypatch.setenv("FPB200_ATTN_V2", "1")
    monkeypatch.setenv("FPB200_ATTN_TPC", str(tpc))
    g = torch.Generator(device="cuda").manual_seed(100 + n_hyp)
    qkv = torch.randn(n_hyp, 400, 1536, device="cuda", generator=g).half()
    out = torch.full((n_hyp, 400, 512), float("nan"), device="cuda", dtype=torch.float16)
    capi.debug_attention(qkv.data_ptr(), out.data_ptr(), n_hyp)
    torch.cuda.synchronize()
    ref = reference(qkv)
    assert torch.isfinite(out).all()
    err = (out.float() - ref).abs().max().item()
    assert err < 2e-3 * 4, err


def test_attention_designs_agree_at_bench_size(monkeypatch):
    """N = 252 (the benchmark's launch): both designs against each other and against fp32."""
    n = 252
    g = torch.Generator(device="cuda").manual_seed(7)
    qkv = torch.randn(n, 400, 1536, device="cuda", generator=g).half()
    outs = []
    for v2 in ("0", "1"):
        monkeypatch.setenv("FPB200_ATTN_V2", v2)
        out = torch.full((n, 400, 512), float("nan"), device="cuda", dtype=torch.float16)
        capi.debug_attention(qkv.data_ptr(), out.data_ptr(), n)
        torch.cuda.synchronize()
        outs.append(out.float())
    assert (outs[0] - outs[1]).abs().max().item() < 2e-3
    for i in range(0, n, 63):
        ref = reference(qkv[i:i + 63])
        assert (outs[0][i:i + 63] - ref).abs().max().item() < 8e-3
