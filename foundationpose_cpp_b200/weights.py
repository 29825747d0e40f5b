"""Weight-file exporter: PyTorch state_dict of the refine / score network -> flat FPB2W file.

The flat file replaces the serialized TensorRT engine the reference loads
(easy_deploy_tool/inference_core/trt_core/src/trt_core.cpp:185-215).  BatchNorm is folded into the
preceding convolution, convolution weights are stored K-major in the layout csrc/conv_gemm.cu
consumes (see layouts.py), everything the tensor cores touch is fp16, everything else fp32.

File layout (little endian):
  char[8]  "FPB2W001"
  u32      kind (0 = refiner, 1 = scorer), u32 tensor count
  per tensor: char[48] name, u32 dtype (0 = f16, 1 = f32), u32 ndim, u32 dims[4], u64 offset, u64 nbytes
  payload, every tensor 256-byte aligned
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np
import torch

from . import layouts

MAGIC = b"FPB2W001"
KIND_REFINER, KIND_SCORER = 0, 1


def _fold(sd: Dict[str, torch.Tensor], conv: str, bn: str, eps: float = 1e-5):
    w, b = sd[conv + ".weight"].float(), sd[conv + ".bias"].float()
    if bn + ".weight" in sd:
        s = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + eps)
        w = w * s[:, None, None, None]
        b = (b - sd[bn + ".running_mean"].float()) * s + sd[bn + ".bias"].float()
    return w, b


# Backbone layer list = DATA (hard part a: the architecture is recalled, not given): every weight
# file carries it as the tensor "arch", f32 [L, 4] rows (section, kind, cout, stride); csrc/network.cu builds its launch
# plan from that table.  section 0 = encodeA (applied to the 2N images), 1 = encodeAB (after the channel concat);
# kind 0 = 7x7 ConvBNReLU on the packed crop (first layer), 1 = 3x3 ConvBNReLU, 2 = ResnetBasicBlock.
SEC_A, SEC_AB = 0, 1
KIND_STEM7, KIND_CBR3, KIND_RES = 0, 1, 2
DEFAULT_ARCH = ((SEC_A, KIND_STEM7, 64, 2), (SEC_A, KIND_CBR3, 128, 2), (SEC_A, KIND_RES, 128, 1), (SEC_A, KIND_RES, 128, 1),
                (SEC_AB, KIND_RES, 256, 1), (SEC_AB, KIND_RES, 256, 1), (SEC_AB, KIND_CBR3, 512, 2),
                (SEC_AB, KIND_RES, 512, 1), (SEC_AB, KIND_RES, 512, 1))


def arch_modules(arch=DEFAULT_ARCH):
    """[(module prefix in the state_dict, file prefix, kind, cout, stride)] - `encodeA.{i}` / `encodeAB.{i}` with i the
    index inside the section, as nn.Sequential numbers them upstream."""
    out, idx = [], {SEC_A: 0, SEC_AB: 0}
    for sec, kind, cout, stride in arch:
        sec, kind = int(sec), int(kind)
        name = "encodeA" if sec == SEC_A else "encodeAB"
        out.append((f"backbone.{name}.{idx[sec]}", f"enc.{'A' if sec == SEC_A else 'AB'}.{idx[sec]}", kind, int(cout), int(stride)))
        idx[sec] += 1
    return out


def arch_from_state_dict(sd) -> tuple:
    """Layer list read off a (canonical) state_dict: `...N.conv.weight` = ConvBNReLU (7x7: the stem; stride 2, as every
    ConvBNReLU of the upstream networks has), `...N.conv1.weight` = ResnetBasicBlock."""
    arch = []
    for sec, name in ((SEC_A, "encodeA"), (SEC_AB, "encodeAB")):
        i = 0
        while True:
            p = f"backbone.{name}.{i}"
            if p + ".conv.weight" in sd:
                w = sd[p + ".conv.weight"]
                arch.append((sec, KIND_STEM7 if w.shape[2] == 7 else KIND_CBR3, int(w.shape[0]), 2))
            elif p + ".conv1.weight" in sd:
                arch.append((sec, KIND_RES, int(sd[p + ".conv1.weight"].shape[0]), 1))
            else:
                break
            i += 1
    if not arch:
        raise KeyError("no backbone.encodeA / encodeAB layers in the state_dict")
    return tuple(arch)


def conv_order(arch=DEFAULT_ARCH):
    """Convolution module names in forward order (ONNX files with anonymous Conv constants are matched by position)."""
    out = []
    for mod, _, kind, _, _ in arch_modules(arch):
        out += [mod + ".conv"] if kind != KIND_RES else [mod + ".conv1", mod + ".conv2"]
    return out


def _backbone_tensors(sd, arch=None):
    arch = arch_from_state_dict(sd) if arch is None else arch
    t = {"arch": torch.tensor([list(map(float, r)) for r in arch], dtype=torch.float32)}
    for mod, f, kind, cout, stride in arch_modules(arch):
        if kind == KIND_RES:
            for c in (1, 2):
                w, b = _fold(sd, f"{mod}.conv{c}", f"{mod}.bn{c}")
                t[f"{f}.c{c}.w"], t[f"{f}.c{c}.b"] = layouts.pack_conv3x3(w).half(), b
        else:
            w, b = _fold(sd, mod + ".conv", mod + ".bn")
            t[f"{f}.w"], t[f"{f}.b"] = (layouts.pack_stem7(w) if kind == KIND_STEM7 else layouts.pack_conv3x3(w)).half(), b
    t["pos.table"] = sd["backbone.pe"].float()[:400]
    return t


def refiner_tensors(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    t = _backbone_tensors(sd)
    for h in ("trans", "rot"):
        q = f"{h}_head.0."
        t[f"{h}.qkv.w"], t[f"{h}.qkv.b"] = sd[q + "self_attn.in_proj_weight"].half(), sd[q + "self_attn.in_proj_bias"].float()
        t[f"{h}.out.w"], t[f"{h}.out.b"] = sd[q + "self_attn.out_proj.weight"].half(), sd[q + "self_attn.out_proj.bias"].float()
        t[f"{h}.ln1.g"], t[f"{h}.ln1.b"] = sd[q + "norm1.weight"].float(), sd[q + "norm1.bias"].float()
        t[f"{h}.ff1.w"], t[f"{h}.ff1.b"] = sd[q + "linear1.weight"].half(), sd[q + "linear1.bias"].float()
        t[f"{h}.ff2.w"], t[f"{h}.ff2.b"] = sd[q + "linear2.weight"].half(), sd[q + "linear2.bias"].float()
        t[f"{h}.ln2.g"], t[f"{h}.ln2.b"] = sd[q + "norm2.weight"].float(), sd[q + "norm2.bias"].float()
        t[f"{h}.head.w"], t[f"{h}.head.b"] = sd[f"{h}_head.1.weight"].float(), sd[f"{h}_head.1.bias"].float()
    return t


def scorer_tensors(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    t = _backbone_tensors(sd)
    t["att.qkv.w"], t["att.qkv.b"] = sd["att.in_proj_weight"].half(), sd["att.in_proj_bias"].float()
    t["att.out.w"], t["att.out.b"] = sd["att.out_proj.weight"].float(), sd["att.out_proj.bias"].float()
    t["cross.qkv.w"], t["cross.qkv.b"] = sd["att_cross.in_proj_weight"].float(), sd["att_cross.in_proj_bias"].float()
    t["cross.out.w"], t["cross.out.b"] = sd["att_cross.out_proj.weight"].float(), sd["att_cross.out_proj.bias"].float()
    t["lin.w"], t["lin.b"] = sd["linear.weight"].float(), sd["linear.bias"].float()
    return t


def write_weight_file(path: str, kind: int, tensors: Dict[str, torch.Tensor]) -> None:
    names = list(tensors)
    header_size = 8 + 8 + len(names) * (48 + 4 + 4 + 16 + 8 + 8)
    off = (header_size + 255) // 256 * 256
    entries, blobs = [], []
    for n in names:
        a = tensors[n].detach().cpu().contiguous()
        assert a.dtype in (torch.float16, torch.float32), (n, a.dtype)
        assert a.dim() <= 4 and len(n) < 48
        raw = a.numpy().tobytes()
        dims = list(a.shape) + [1] * (4 - a.dim())
        entries.append(struct.pack("<48sII4IQQ", n.encode(), 0 if a.dtype == torch.float16 else 1, a.dim(), *dims,
                                   off, len(raw)))
        blobs.append((off, raw))
        off = (off + len(raw) + 255) // 256 * 256
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<II", kind, len(names)))
        for e in entries:
            f.write(e)
        for o, raw in blobs:
            f.seek(o)
            f.write(raw)
        f.truncate(off)


def export_refiner(path: str, state_dict) -> None:
    write_weight_file(path, KIND_REFINER, refiner_tensors(state_dict))


def export_scorer(path: str, state_dict) -> None:
    write_weight_file(path, KIND_SCORER, scorer_tensors(state_dict))


def read_weight_file(path: str) -> Dict[str, np.ndarray]:
    """Reader (tests / tooling)."""
    out = {}
    with open(path, "rb") as f:
        buf = f.read()
    assert buf[:8] == MAGIC
    _, n = struct.unpack_from("<II", buf, 8)
    p = 16
    for _ in range(n):
        name, dt, nd, d0, d1, d2, d3, off, nb = struct.unpack_from("<48sII4IQQ", buf, p)
        p += 48 + 4 + 4 + 16 + 8 + 8
        arr = np.frombuffer(buf, dtype=np.float16 if dt == 0 else np.float32, count=nb // (2 if dt == 0 else 4),
                            offset=off).reshape([d0, d1, d2, d3][:nd])
        out[name.rstrip(b"\0").decode()] = arr
    return out


# ----------------------------------------------------------------------------- seeded weights
def positional_table(max_len: int = 400, d: int = 512) -> torch.Tensor:
    """PositionalEmbedding of the upstream networks: pe[p,2i] = sin(p * e^(-2i ln(1e4)/d)), pe[p,2i+1] = cos."""
    import math
    pe = torch.zeros(max_len, d)
    pos = torch.arange(0, max_len, dtype=torch.float32).unsqueeze(1)
    div = torch.exp(torch.arange(0, d, 2, dtype=torch.float32) * (-math.log(10000.0) / d))
    pe[:, 0::2] = torch.sin(pos * div)
    pe[:, 1::2] = torch.cos(pos * div)
    return pe


def seeded_state_dict(kind: int, seed: int, arch=DEFAULT_ARCH) -> Dict[str, torch.Tensor]:
    """Random-init state_dict with the upstream parameter names and shapes (no checkpoint is
    reachable offline): He-normal convolutions, non-trivial BatchNorm statistics, small transformer
    projections.  Used by bench.py; the tests export the oracle modules' own state_dict instead."""
    import math
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k):
        sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / (cin * k * k))
        sd[name + ".bias"] = torch.randn(cout, generator=g) * 0.05

    def bn(name, c):
        sd[name + ".weight"] = 0.6 + 0.3 * torch.rand(c, generator=g)
        sd[name + ".bias"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn(c, generator=g) * 0.1
        sd[name + ".running_var"] = 0.8 + 0.4 * torch.rand(c, generator=g)

    def linear(name, o, i, gain=1.0):
        sd[name + ".weight"] = torch.randn(o, i, generator=g) * (gain / math.sqrt(i))
        sd[name + ".bias"] = torch.randn(o, generator=g) * 0.02

    def cbr(name, cout, cin, k):
        conv(name + ".conv", cout, cin, k)
        bn(name + ".bn", cout)

    def res(name, c):
        for i in (1, 2):
            conv(f"{name}.conv{i}", c, c, 3)
            bn(f"{name}.bn{i}", c)

    def mha(name):
        sd[name + ".in_proj_weight"] = torch.randn(1536, 512, generator=g) * 0.04
        sd[name + ".in_proj_bias"] = torch.randn(1536, generator=g) * 0.02
        linear(name + ".out_proj", 512, 512)

    b = "backbone."
    cin = {SEC_A: 6, SEC_AB: None}
    for mod, _, k, cout, _ in arch_modules(arch):
        sec = SEC_A if ".encodeA." in mod else SEC_AB
        if cin[sec] is None:
            cin[sec] = 2 * cin[SEC_A]  # channel concat of the two branches
        if k == KIND_RES:
            res(mod, cout)
        else:
            cbr(mod, cout, cin[sec], 7 if k == KIND_STEM7 else 3)
        cin[sec] = cout
    sd[b + "pe"] = positional_table()
    if kind == KIND_REFINER:
        for h in ("trans_head", "rot_head"):
            mha(f"{h}.0.self_attn")
            linear(f"{h}.0.linear1", 512, 512)
            linear(f"{h}.0.linear2", 512, 512)
            for nrm in ("norm1", "norm2"):
                sd[f"{h}.0.{nrm}.weight"] = 1.0 + 0.1 * torch.randn(512, generator=g)
                sd[f"{h}.0.{nrm}.bias"] = 0.05 * torch.randn(512, generator=g)
            linear(f"{h}.1", 3, 512, gain=0.3)
    else:
        mha("att")
        mha("att_cross")
        linear("linear", 1, 512, gain=4.0)
    return sd


def write_seeded(dirpath: str, refiner_seed: int = 2, scorer_seed: int = 3):
    """Writes refiner.fpb2w / scorer.fpb2w with seeded weights into dirpath; returns the two paths."""
    import os
    os.makedirs(dirpath, exist_ok=True)
    rp, sp = os.path.join(dirpath, "refiner.fpb2w"), os.path.join(dirpath, "scorer.fpb2w")
    export_refiner(rp, seeded_state_dict(KIND_REFINER, refiner_seed))
    export_scorer(sp, seeded_state_dict(KIND_SCORER, scorer_seed))
    return rp, sp


# ----------------------------------------------------------------------------- importers
# conv layers in forward order; ONNX exporters that fold BatchNorm leave the convolutions with
# anonymous initializer names ("onnx::Conv_123"), so they are matched by position
_CONV_ORDER = conv_order(DEFAULT_ARCH)


def canonical_state_dict(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Maps the parameter names of an upstream NVlabs/FoundationPose checkpoint onto the names this module uses
    (those of oracle/nets.py).  Upstream names, as recalled from learning/models/{refine_network,score_network,
    network_modules}.py (no checkpoint can be fetched in this environment to confirm them):
      * no `backbone.` prefix: `encodeA.*` / `encodeAB.*` (refiner), `encoderA.*` / `encoderAB.*` (scorer);
      * ConvBNReLU keeps its layers in `self.net = nn.Sequential(conv, bn, relu)`: `encodeA.0.net.0.*` is the
        convolution, `encodeA.0.net.1.*` the BatchNorm; ResnetBasicBlock uses `conv1 / bn1 / conv2 / bn2`;
      * the positional table is the buffer `pos_embed.pe` of shape [1, 400, 512];
      * a `module.` prefix (DataParallel) may precede everything.
    Names that already are canonical pass through, so the function is idempotent."""
    import re
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[7:]
        if k.endswith("num_batches_tracked"):
            continue
        k = re.sub(r"^(backbone\.)?encoder(AB?)\.", r"\1encode\2.", k)  # scorer spelling
        if re.match(r"^encode(AB?)\.", k):
            k = "backbone." + k
        k = re.sub(r"^(backbone\.encodeAB?\.\d+)\.net\.0\.", r"\1.conv.", k)
        k = re.sub(r"^(backbone\.encodeAB?\.\d+)\.net\.1\.", r"\1.bn.", k)
        if k in ("pos_embed.pe", "backbone.pos_embed.pe"):
            k = "backbone.pe"
        if k == "backbone.pe":
            v = v.reshape(-1, v.shape[-1])  # [1, 400, 512] -> [400, 512]
        out[k] = v
    return out


def load_checkpoint(path: str) -> Dict[str, torch.Tensor]:
    """PyTorch checkpoint (.pth / .pt) -> flat state_dict with the upstream parameter names.
    Accepts a bare state_dict or the usual wrappers ({'model': ...}, {'state_dict': ...}) and strips a
    DataParallel 'module.' prefix."""
    obj = torch.load(path, map_location="cpu", weights_only=True)
    for key in ("model", "state_dict", "model_state_dict"):
        if isinstance(obj, dict) and key in obj and isinstance(obj[key], dict):
            obj = obj[key]
    if not isinstance(obj, dict) or not all(isinstance(v, torch.Tensor) for v in obj.values()):
        raise ValueError(f"{path}: not a state_dict")
    return canonical_state_dict(obj)


# dense layers in forward order with their torch shapes [out, in]: ONNX exporters turn nn.Linear / the packed MHA
# in-projection applied to 3-D activations into MatMul(x, W^T) + Add(bias) with anonymous, transposed constants
_DENSE_ORDER = {
    KIND_REFINER: [(f"{h}_head.0.{n}", o, i) for h in ("trans", "rot") for n, o, i in
                   (("self_attn.in_proj", 1536, 512), ("self_attn.out_proj", 512, 512), ("linear1", 512, 512),
                    ("linear2", 512, 512))] ,
    KIND_SCORER: [("att.in_proj", 1536, 512), ("att.out_proj", 512, 512), ("att_cross.in_proj", 1536, 512),
                  ("att_cross.out_proj", 512, 512), ("linear", 1, 512)],
}
_DENSE_ORDER[KIND_REFINER] = (_DENSE_ORDER[KIND_REFINER][:4] + [("trans_head.1", 3, 512)] +
                              _DENSE_ORDER[KIND_REFINER][4:] + [("rot_head.1", 3, 512)])


def _dense_key(name: str, what: str) -> str:
    return name + ("_weight" if what == "weight" else "_bias") if name.endswith("in_proj") else f"{name}.{what}"


def _recover_dense(sd, inits, nodes, kind, path):
    """Fills the dense-layer weights (and biases) of `sd` that the ONNX file holds only as anonymous MatMul / Gemm
    constants, matched by position in graph order against _DENSE_ORDER; MatMul constants are transposed back to
    torch's [out, in]; an in-projection exported as three 512-wide products (q, k, v slices) is concatenated back."""
    want = _DENSE_ORDER[kind]
    if all(_dense_key(n, "weight") in sd for n, _, _ in want):
        return
    consumers = {}
    for idx, (op, ins, outs) in enumerate(nodes):
        for i in ins:
            consumers.setdefault(i, []).append(idx)
    found = []  # (weight [out, in], bias or None)
    for op, ins, outs in nodes:
        if op not in ("MatMul", "Gemm"):
            continue
        const = [i for i in ins if i in inits and inits[i].ndim == 2]
        if len(const) != 1:
            continue  # activation x activation products of the attention itself
        w = torch.from_numpy(np.asarray(inits[const[0]])).float()
        w = w.t() if op == "MatMul" else w  # Gemm as exported from nn.Linear carries transB = 1
        bias = None
        if op == "Gemm":
            b = [i for i in ins if i in inits and inits[i].ndim == 1]
            bias = torch.from_numpy(np.asarray(inits[b[0]])).float() if b else None
        else:
            for c in consumers.get(outs[0], []):
                cop, cins, _ = nodes[c]
                b = [i for i in cins if i in inits and inits[i].ndim == 1]
                if cop == "Add" and b:
                    bias = torch.from_numpy(np.asarray(inits[b[0]])).float()
                    break
        found.append((w.contiguous(), bias))
    k = 0
    for name, o, i in want:
        if k < len(found) and tuple(found[k][0].shape) == (o, i):
            w, b = found[k]
            k += 1
        elif o == 1536 and k + 2 < len(found) and all(tuple(found[k + j][0].shape) == (512, i) for j in range(3)):
            w = torch.cat([found[k + j][0] for j in range(3)], 0)
            bs = [found[k + j][1] for j in range(3)]
            b = torch.cat(bs, 0) if all(x is not None for x in bs) else None
            k += 3
        else:
            got = tuple(found[k][0].shape) if k < len(found) else None
            raise ValueError(f"{path}: dense layer '{name}' expects a [{o},{i}] weight at MatMul / Gemm #{k}, found {got}")
        sd.setdefault(_dense_key(name, "weight"), w)
        if b is not None:
            sd.setdefault(_dense_key(name, "bias"), b)
    if k != len(found):
        raise ValueError(f"{path}: {len(found) - k} dense layers with constant weights are left over after matching")


def onnx_state_dict(path: str, kind: int = -1) -> Dict[str, torch.Tensor]:
    """ONNX model (the refiner_hwc.onnx / scorer_hwc.onnx the reference feeds to trtexec,
    tools/cvt_onnx2trt.bash:3-15) -> state_dict with the upstream parameter names.

    Named initializers are taken as they are.  Convolutions whose weights lost their names because the
    exporter folded BatchNorm into them are matched BY POSITION: distinct Conv weight initializers in
    graph order = the 15 backbone convolutions in forward order (the siamese encodeA branch shares one
    set of weights, so its second appearance is skipped).  Dense layers whose weights became anonymous
    transposed MatMul constants are matched by position too (_recover_dense; needs `kind`).  Anything
    else that is missing is reported by the exporter as a KeyError naming the parameter."""
    from . import onnx_reader
    inits, nodes = onnx_reader.read_onnx(path)
    sd = canonical_state_dict({k: torch.from_numpy(v) for k, v in inits.items() if v.dtype in (np.float32, np.float16)})
    if _CONV_ORDER[0] + ".weight" not in sd:
        seen, convs = set(), []
        for op, ins, _ in nodes:
            if op == "Conv" and len(ins) >= 2 and ins[1] in inits and ins[1] not in seen:
                seen.add(ins[1])
                convs.append(ins)
        if len(convs) != len(_CONV_ORDER):
            raise ValueError(f"{path}: expected {len(_CONV_ORDER)} distinct convolutions, found {len(convs)}")
        for name, ins in zip(_CONV_ORDER, convs):
            w = torch.from_numpy(inits[ins[1]]).float()
            sd[name + ".weight"] = w
            sd[name + ".bias"] = (torch.from_numpy(inits[ins[2]]).float() if len(ins) > 2 and ins[2] in inits
                                  else torch.zeros(w.shape[0]))
    if kind in _DENSE_ORDER:
        _recover_dense(sd, inits, nodes, kind, path)
    if "backbone.pe" not in sd or tuple(sd["backbone.pe"].shape) != (400, 512):
        sd["backbone.pe"] = positional_table()  # a pure function of the architecture, not a trained tensor
    return {k: v.float() for k, v in sd.items()}


def import_weights(src: str, dst: str, kind: int) -> None:
    """src: .pth/.pt checkpoint or .onnx model of the refine (kind 0) / score (kind 1) network ->
    dst: flat FPB2W file for fpb200_create."""
    sd = onnx_state_dict(src, kind) if src.lower().endswith(".onnx") else load_checkpoint(src)
    tensors = refiner_tensors(sd) if kind == KIND_REFINER else scorer_tensors(sd)
    write_weight_file(dst, kind, tensors)
