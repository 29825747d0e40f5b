"""fp32 operators of the scorer tail (csrc/net_ops.cu: small_linear_kernel, row_dot_kernel,
cross_attention_kernel) against plain PyTorch fp32 references of the same ops.  In the reference these run
inside the scorer TensorRT engine (trt_core.cpp:367); upstream: ScoreNetMultiPair's cross-hypothesis
MultiheadAttention + Linear(512, 1)."""
import pytest
import torch

from foundationpose_cpp_b200 import capi

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,k,outs,bias", [(252, 512, 512, True), (252, 512, 1536, True), (252, 512, 1, True),
                                             (1, 512, 3, True), (33, 36, 65, False), (7, 512, 1, False)])
def test_small_linear(rows, k, outs, bias):
    g = torch.Generator(device="cuda").manual_seed(rows * 7 + outs)
    x = torch.randn(rows, k, device="cuda", generator=g)
    w = torch.randn(outs, k, device="cuda", generator=g) / k ** 0.5
    b = torch.randn(outs, device="cuda", generator=g) if bias else None
    out = torch.full((rows, outs), float("nan"), device="cuda")
    capi.debug_small_linear(x.data_ptr(), w.data_ptr(), b.data_ptr() if bias else 0, out.data_ptr(), rows, k, outs)
    torch.cuda.synchronize()
    ref = (x.double() @ w.double().T + (b.double() if bias else 0)).float()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("n", [1, 5, 32, 33, 252, 300])
def test_cross_attention(n):
    g = torch.Generator(device="cuda").manual_seed(n)
    qkv = torch.randn(n, 1536, device="cuda", generator=g) * 1.5
    out = torch.full((n, 512), float("nan"), device="cuda")
    capi.debug_cross_attention(qkv.data_ptr(), out.data_ptr(), n)
    torch.cuda.synchronize()
    q, k, v = [t.double().view(n, 4, 128).transpose(0, 1) for t in qkv.split(512, dim=-1)]
    s = (q @ k.transpose(-1, -2)) * (128 ** -0.5)
    ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(n, 512).float()
    assert torch.isfinite(out).all()
    assert (out - ref).abs().max().item() < 2e-5 * max(1.0, ref.abs().max().item())
