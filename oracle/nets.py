"""ORACLE — test infrastructure only.  PyTorch fp32 restatement of the refine and score networks.

The reference does not contain these networks: it loads two opaque TensorRT engines built from
refiner_hwc.onnx / scorer_hwc.onnx (tools/cvt_onnx2trt.bash:3-15; README.md:72, files git-ignored).
What the reference pins is only the I/O contract (detection_6d_foundationpose/src/foundationpose.cpp:78-83,
349-354; simple_tests/src/test_foundationpose.cpp:29,35): inputs render_input / transf_input
f32 NHWC [N,160,160,6], outputs trans [N,3], rot [N,3] (refiner) and scores [N,1] (scorer).
The layer structure below is the published NVlabs/FoundationPose architecture
(learning/models/refine_network.py, score_network.py, network_modules.py) as recalled in SURVEY.md
§2.3 — "parity unpinned": no ONNX / checkpoint is available in this environment to confirm it,
so the backbone's layer list is DATA on both sides: `Backbone(arch)` here, the "arch" tensor of the weight file
(foundationpose_cpp_b200/weights.py: DEFAULT_ARCH, arch_from_state_dict) from which csrc/network.cu builds its plan.
The transformer heads (400 tokens x 512, 4 heads) are fixed in the kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F


class ConvBNReLU(nn.Module):
    def __init__(self, cin, cout, k, stride):
        super().__init__()
        self.conv = nn.Conv2d(cin, cout, k, stride, (k - 1) // 2, bias=True)
        self.bn = nn.BatchNorm2d(cout)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)))


class ResnetBasicBlock(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1, bias=True)
        self.bn1 = nn.BatchNorm2d(c)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1, bias=True)
        self.bn2 = nn.BatchNorm2d(c)

    def forward(self, x):
        out = F.relu(self.bn1(self.conv1(x)))
        out = self.bn2(self.conv2(out))
        return F.relu(out + x)


def positional_table(max_len=400, d=512):
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


# (section, kind, cout, stride): section 0 = encodeA, 1 = encodeAB; kind 0 = 7x7 ConvBNReLU, 1 = 3x3 ConvBNReLU,
# 2 = ResnetBasicBlock.  The same table as foundationpose_cpp_b200.weights.DEFAULT_ARCH (kept separate: the oracle does
# not import the product).
DEFAULT_ARCH = ((0, 0, 64, 2), (0, 1, 128, 2), (0, 2, 128, 1), (0, 2, 128, 1),
                (1, 2, 256, 1), (1, 2, 256, 1), (1, 1, 512, 2), (1, 2, 512, 1), (1, 2, 512, 1))


class Backbone(nn.Module):
    """encodeA on cat([A,B]) then encodeAB on the channel concat; returns tokens [N,400,512] + PE."""

    def __init__(self, arch=DEFAULT_ARCH):
        super().__init__()
        secs, cin = {0: [], 1: []}, {0: 6, 1: None}
        for sec, kind, cout, stride in arch:
            if cin[sec] is None:
                cin[sec] = 2 * cin[0]
            if kind == 2:
                assert cout == cin[sec], "a ResnetBasicBlock keeps its channel count"
                secs[sec].append(ResnetBasicBlock(cout))
            else:
                secs[sec].append(ConvBNReLU(cin[sec], cout, 7 if kind == 0 else 3, stride))
            cin[sec] = cout
        self.encodeA = nn.Sequential(*secs[0])
        self.encodeAB = nn.Sequential(*secs[1])
        self.register_buffer("pe", positional_table())

    def forward(self, a, b):
        n = a.shape[0]
        x = self.encodeA(torch.cat([a, b], 0))
        ab = torch.cat([x[:n], x[n:]], 1)
        ab = self.encodeAB(ab)
        tok = ab.reshape(n, 512, -1).permute(0, 2, 1)
        return tok + self.pe[None, : tok.shape[1]]


class RefineNet(nn.Module):
    def __init__(self, arch=DEFAULT_ARCH):
        super().__init__()
        self.backbone = Backbone(arch)
        self.trans_head = nn.Sequential(nn.TransformerEncoderLayer(512, 4, 512, dropout=0.0, batch_first=True),
                                        nn.Linear(512, 3))
        self.rot_head = nn.Sequential(nn.TransformerEncoderLayer(512, 4, 512, dropout=0.0, batch_first=True),
                                      nn.Linear(512, 3))

    def forward(self, a, b):
        """a, b: [N,6,160,160] (NCHW views of the reference's NHWC blobs)."""
        tok = self.backbone(a, b)
        return self.trans_head(tok).mean(1), self.rot_head(tok).mean(1)


class ScoreNet(nn.Module):
    def __init__(self, arch=DEFAULT_ARCH):
        super().__init__()
        self.backbone = Backbone(arch)
        self.att = nn.MultiheadAttention(512, 4, batch_first=True, bias=True)
        self.att_cross = nn.MultiheadAttention(512, 4, batch_first=True, bias=True)
        self.linear = nn.Linear(512, 1)

    def features(self, a, b):
        tok = self.backbone(a, b)
        out, _ = self.att(tok, tok, tok, need_weights=False)
        return out.mean(1)

    def tail(self, feats):
        x = feats[None]
        x, _ = self.att_cross(x, x, x, need_weights=False)
        return self.linear(x).reshape(-1)

    def forward(self, a, b):
        return self.tail(self.features(a, b))


def seeded_init(net: nn.Module, seed: int) -> nn.Module:
    """Seeded weights (no checkpoint is reachable offline): He-normal convs, non-trivial BatchNorm
    statistics, small transformer projections so softmax logits stay O(1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, m in net.named_modules():
            if isinstance(m, nn.Conv2d):
                fan_in = m.in_channels * m.kernel_size[0] * m.kernel_size[1]
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * math.sqrt(2.0 / fan_in))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.copy_(0.6 + 0.3 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
                m.running_mean.copy_(torch.randn(m.running_mean.shape, generator=g) * 0.1)
                m.running_var.copy_(0.8 + 0.4 * torch.rand(m.running_var.shape, generator=g))
            elif isinstance(m, nn.Linear):
                gain = {3: 0.3, 1: 4.0}.get(m.out_features, 1.0)  # pose heads small, score head wide
                m.weight.copy_(torch.randn(m.weight.shape, generator=g) * (gain / math.sqrt(m.in_features)))
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.02)
            elif isinstance(m, nn.MultiheadAttention):
                m.in_proj_weight.copy_(torch.randn(m.in_proj_weight.shape, generator=g) * 0.04)
                m.in_proj_bias.copy_(torch.randn(m.in_proj_bias.shape, generator=g) * 0.02)
            elif isinstance(m, nn.LayerNorm):
                m.weight.copy_(1.0 + 0.1 * torch.randn(m.weight.shape, generator=g))
                m.bias.copy_(0.05 * torch.randn(m.bias.shape, generator=g))
    net.eval()
    return net


def nhwc_to_inputs(x6: torch.Tensor) -> torch.Tensor:
    """[N,160,160,6] NHWC blob -> NCHW tensor for the modules above."""
    return x6.permute(0, 3, 1, 2).contiguous()
