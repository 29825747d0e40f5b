"""CPU tests: C-ABI surface, weight-file round trip, layouts, sharding arithmetic (gloo, world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from foundationpose_cpp_b200 import build, capi
    build.build()
    hdr = open(os.path.join(ROOT, "include", "fpb200.h")).read()
    names = sorted(set(re.findall(r"\b(fpb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = capi.lib()
    for n in names:
        assert hasattr(lib, n), f"libfpb200.so does not export {n}"
    # no compute without a GPU: option defaults only
    o = capi.Options()
    lib.fpb200_default_options(ctypes.byref(o))
    assert (o.max_hyps, o.shard_world, o.shard_rank) == (252, 1, 0)


def test_create_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from foundationpose_cpp_b200 import api, synth, weights
    with tempfile.TemporaryDirectory() as d:
        rp, sp = weights.write_seeded(d)
        mesh = synth.uv_ellipsoid("m", slices=8, stacks=5, tex_size=16)
        with pytest.raises(Exception) as e:
            api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, 480, 640)
        assert "no CPU fallback" in str(e.value) or "CUDA" in str(e.value)


def test_weight_file_roundtrip_and_bn_fold():
    from foundationpose_cpp_b200 import layouts, weights
    sd = weights.seeded_state_dict(weights.KIND_REFINER, 7)
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "r.fpb2w")
        weights.export_refiner(p, sd)
        t = weights.read_weight_file(p)
    assert t["enc.A.0.w"].shape == (64, 448) and t["enc.A.0.w"].dtype == np.float16
    assert t["enc.AB.4.c2.w"].shape == (512, 4608)
    assert t["trans.qkv.w"].shape == (1536, 512) and t["rot.head.w"].shape == (3, 512)
    assert t["pos.table"].shape == (400, 512)
    # folded conv == conv + BN(eval) on random input
    x = torch.randn(2, 64, 12, 12)
    w, b = sd["backbone.encodeA.1.conv.weight"], sd["backbone.encodeA.1.conv.bias"]
    y = F.conv2d(x, w, b, stride=2, padding=1)
    y = F.batch_norm(y, sd["backbone.encodeA.1.bn.running_mean"], sd["backbone.encodeA.1.bn.running_var"],
                     sd["backbone.encodeA.1.bn.weight"], sd["backbone.encodeA.1.bn.bias"], False, 0.0, 1e-5)
    wf = torch.from_numpy(t["enc.A.1.w"].astype(np.float32)).reshape(128, 3, 3, 64).permute(0, 3, 1, 2)
    y2 = F.conv2d(x, wf, torch.from_numpy(t["enc.A.1.b"]), stride=2, padding=1)
    assert torch.allclose(y, y2, atol=2e-2, rtol=2e-2)  # fp16 weight rounding


def test_stem_layout_is_a_valid_gemm_view():
    """The overlapping-stride K view of the padded NHWC8 input times pack_stem7 weights == conv2d 7x7/s2."""
    from foundationpose_cpp_b200 import layouts
    torch.manual_seed(0)
    x = torch.randn(1, 160, 160, 6)
    w = torch.randn(4, 6, 7, 7)
    buf = layouts.pad_stem_input(x).float().numpy()
    wp = layouts.pack_stem7(w).numpy()  # [4, 7*64]
    pitch = layouts.STEM_PITCH * 8
    ref = F.conv2d(x.half().float().permute(0, 3, 1, 2), w, stride=2, padding=3)[0].permute(1, 2, 0).numpy()
    for (oy, ox) in [(0, 0), (0, 79), (79, 0), (79, 79), (40, 13)]:
        acc = np.zeros(4, np.float64)
        for ky in range(7):
            iy = 2 * oy + ky - 3
            if iy < 0 or iy >= 160:
                continue  # TMA out-of-bounds rows are zero-filled
            start = iy * pitch + ox * 16  # 32-byte stride per output x
            kvec = buf[start:start + 64]
            acc += wp[:, ky * 64:(ky + 1) * 64].astype(np.float64) @ kvec
        np.testing.assert_allclose(acc, ref[oy, ox], rtol=1e-3, atol=1e-3)


def test_phase_split_roundtrip():
    from foundationpose_cpp_b200 import layouts
    x = torch.arange(2 * 8 * 8 * 3, dtype=torch.float32).reshape(2, 8, 8, 3)
    s = layouts.phase_split(x)
    assert s.shape == (2, 4, 4, 4, 3)
    assert torch.equal(layouts.phase_merge(s), x)
    assert s[0, 3, 1, 2, 0] == x[0, 3, 5, 0]  # plane 3 = odd row, odd col


def test_shard_ranges():
    from foundationpose_cpp_b200 import sharding
    sizes = [sharding.shard_range(252, r, 8) for r in range(8)]
    assert [b - a for a, b in sizes] == [32, 32, 32, 32, 31, 31, 31, 31]
    assert sizes[0][0] == 0 and sizes[-1][1] == 252
    for w in (1, 2, 3, 4, 8):
        rows = [sharding.padded_row(252, w, i) for i in range(252)]
        cap = sharding.shard_capacity(252, w)
        for r in range(w):
            lo, hi = sharding.shard_range(252, r, w)
            assert rows[lo:hi] == list(range(r * cap, r * cap + hi - lo))
    assert sharding.shard_range(1, 0, 1) == (0, 1)


_GLOO_WORKER = r"""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from foundationpose_cpp_b200 import sharding
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n = 13
cap = sharding.shard_capacity(n, world)
lo, hi = sharding.shard_range(n, rank, world)
local = torch.full((cap, sharding.EXCH_LD), -1.0)
local[: hi - lo] = torch.arange(lo, hi, dtype=torch.float32)[:, None]
allbuf = torch.empty(world * cap, sharding.EXCH_LD)
dist.all_gather_into_tensor(allbuf, local)
dense = torch.stack([allbuf[sharding.padded_row(n, world, i)] for i in range(n)])
assert torch.equal(dense[:, 0], torch.arange(n, dtype=torch.float32)), dense[:, 0]
# exchange-handle plumbing of the peer-memory exchange (fpb200_exchange_export -> gather -> fpb200_exchange_connect)
handles = sharding.gather_handles(bytes([rank + 1] * 64), dist)
assert handles == [bytes([r + 1] * 64) for r in range(world)], handles
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_allgather_layout_gloo_world2():
    with tempfile.TemporaryDirectory() as d:
        script = os.path.join(d, "w.py")
        open(script, "w").write(_GLOO_WORKER)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                            "--master-addr", "127.0.0.1", "--master-port", "29611", script, ROOT],
                           capture_output=True, text=True, timeout=240)
        assert r.returncode == 0, r.stdout + r.stderr
        assert r.stdout.count("ok") == 2


def test_obj_mesh_loader_roundtrip(tmp_path):
    from foundationpose_cpp_b200 import mesh_loader, synth
    m = synth.uv_ellipsoid("obj", slices=16, stacks=9, tex_size=32, seed=3)
    p = str(tmp_path / "m.obj")
    mesh_loader.save_obj(p, m.vertices, m.normals, m.texcoords, m.faces, m.texture)
    ld = mesh_loader.CreateAssimpMeshLoader("obj", p)
    assert ld.GetName() == "obj"
    # vertices are re-indexed in order of first use (like assimp's JoinIdenticalVertices): compare per corner
    f2 = ld.GetMeshTriangleFaces()
    assert f2.shape == m.faces.shape and ld.GetMeshNumVertices() <= len(m.vertices)
    np.testing.assert_allclose(ld.GetMeshVertices()[f2], m.vertices[m.faces], atol=1e-6)
    np.testing.assert_allclose(ld.GetMeshVertexNormals()[f2], m.normals[m.faces], atol=1e-5)
    np.testing.assert_allclose(ld.GetMeshTextureCoords()[f2], m.texcoords[m.faces], atol=1e-6)
    assert np.array_equal(ld.GetTextureMap(), m.texture)
    np.testing.assert_allclose(ld.GetMeshModelCenter(), m.center, atol=1e-6)
    assert abs(ld.GetMeshDiameter() - m.diameter) < 1e-3  # max pairwise distance of the tessellated ellipsoid
    # quads are fan-triangulated, shared corners are joined, missing texture -> 2x2 grey like the reference
    q = str(tmp_path / "quad.obj")
    open(q, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nf 1/1 2/2 3/3 4/4\n")
    v, n, t, f, tex = mesh_loader.load_obj(q)
    assert f.tolist() == [[0, 1, 2], [0, 2, 3]] and v.shape == (4, 3)
    np.testing.assert_allclose(n, [[0, 0, 1]] * 4)
    assert tex.shape == (2, 2, 3) and np.all(tex == 100)
    open(q, "w").write("v 0 0 0\nv 1 0 0\nv 1 1 0\nf 1 2 3\n")
    with pytest.raises(ValueError):
        mesh_loader.load_obj(q)


def test_orient_bounds_and_bbox_pose_known_answer():
    """GetOrientBounds / GetObjectDimension (ComputeOBB, assimp_mesh_loader.cpp:62-113) and ConvertPoseMesh2BBox
    (mesh_loader.hpp:75-81: pose * T(-center) * orient_bounds) on a point cloud whose answer is known: a 0.1 x 0.2 x 0.4
    box grid, rotated and shifted.  PCA axes come in ascending-variance order; signs are the solver's."""
    from foundationpose_cpp_b200 import api
    g = np.stack(np.meshgrid(np.linspace(-0.05, 0.05, 5), np.linspace(-0.1, 0.1, 7), np.linspace(-0.2, 0.2, 9),
                             indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    c, s_ = np.cos(0.3), np.sin(0.3)
    R = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]], np.float32) @ np.array([[1, 0, 0], [0, c, -s_], [0, s_, c]], np.float32)
    shift = np.array([0.3, -0.1, 0.05], np.float32)
    v = g @ R.T + shift
    faces = np.array([[0, 1, 2]], np.int32)
    ld = api.ArrayMeshLoader("box", v, np.tile([0, 0, 1.0], (len(v), 1)), np.zeros((len(v), 2)), faces,
                             np.zeros((2, 2, 3), np.uint8))
    ob, dim = ld.GetOrientBounds(), ld.GetObjectDimension()
    np.testing.assert_allclose(dim, [0.1, 0.2, 0.4], atol=1e-5)
    np.testing.assert_allclose(ob[:3, 3], shift, atol=1e-6)  # vertex mean
    np.testing.assert_allclose(np.abs(ob[:3, :3].T @ R), np.eye(3), atol=1e-4)  # the box axes, up to sign
    np.testing.assert_allclose(ob[3], [0, 0, 0, 1])
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = R.T
    pose[:3, 3] = [0.0, 0.1, 0.7]
    box = api.ConvertPoseMesh2BBox(pose, ld)
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3] = -ld.GetMeshModelCenter()
    np.testing.assert_allclose(box, pose @ tf @ ob, atol=1e-6)
    # the centre is the AABB midpoint (assimp_mesh_loader.cpp:179-180), the box origin the vertex mean: for this
    # symmetric cloud both are `shift`, so the box pose keeps the translation of the centred-mesh pose
    np.testing.assert_allclose(box[:3, 3], pose[:3, 3], atol=1e-5)


def test_weight_importers_pth_and_onnx(tmp_path):
    """a .pth checkpoint and an ONNX model (named initializers, and the BN-folded form
    with anonymous Conv weights matched by position) produce the same FPB2W file as the exporter."""
    import numpy as np
    import torch
    from foundationpose_cpp_b200 import onnx_reader, weights
    for kind in (weights.KIND_REFINER, weights.KIND_SCORER):
        sd = weights.seeded_state_dict(kind, 11 + kind)
        ref = tmp_path / f"ref{kind}.fpb2w"
        (weights.export_refiner if kind == weights.KIND_REFINER else weights.export_scorer)(str(ref), sd)
        want = weights.read_weight_file(str(ref))
        # 1) .pth, wrapped and with a DataParallel prefix
        pth = tmp_path / f"k{kind}.pth"
        torch.save({"model": {"module." + k: v for k, v in sd.items()}}, pth)
        out = tmp_path / f"pth{kind}.fpb2w"
        weights.import_weights(str(pth), str(out), kind)
        got = weights.read_weight_file(str(out))
        assert got.keys() == want.keys() and all(np.array_equal(got[k], want[k]) for k in want)
        # 2) ONNX with named initializers (fp16 and fp32 tensors, raw_data encoding)
        onx = tmp_path / f"k{kind}.onnx"
        onnx_reader.write_onnx(str(onx), {k: v.numpy() for k, v in sd.items()})
        inits, _ = onnx_reader.read_onnx(str(onx))
        assert all(np.array_equal(inits[k], sd[k].numpy()) for k in sd)
        out = tmp_path / f"onnx{kind}.fpb2w"
        weights.import_weights(str(onx), str(out), kind)
        got = weights.read_weight_file(str(out))
        assert all(np.array_equal(got[k], want[k]) for k in want)
        # 3) ONNX as exporters write it in eval mode: BN folded, Conv weights anonymous, siamese branch twice
        folded, nodes, n = {}, [], 0
        for name in weights._CONV_ORDER:
            bn = name.replace(".conv", ".bn") if name.endswith(".conv") else name.replace("conv", "bn")
            w, b = weights._fold(sd, name, bn)
            folded[f"onnx::Conv_{n}"], folded[f"onnx::Conv_{n + 1}"] = w.numpy(), b.numpy()
            nodes.append(("Conv", [f"x{n}", f"onnx::Conv_{n}", f"onnx::Conv_{n + 1}"], [f"y{n}"]))
            if "encodeA." in name:  # second siamese branch reuses the initializers
                nodes.append(("Conv", [f"xb{n}", f"onnx::Conv_{n}", f"onnx::Conv_{n + 1}"], [f"yb{n}"]))
            n += 2
        for k, v in sd.items():
            if not k.startswith("backbone.encode"):
                folded[k] = v.numpy()
        onx2 = tmp_path / f"folded{kind}.onnx"
        onnx_reader.write_onnx(str(onx2), folded, nodes)
        out = tmp_path / f"folded{kind}.fpb2w"
        weights.import_weights(str(onx2), str(out), kind)
        got = weights.read_weight_file(str(out))
        for k in want:
            np.testing.assert_allclose(got[k].astype(np.float32), want[k].astype(np.float32), rtol=2e-3, atol=1e-6)


def _upstream_names(sd, kind):
    """Canonical (oracle/nets.py) names -> the names of an upstream NVlabs/FoundationPose checkpoint as recalled
    (weights.canonical_state_dict documents them): no backbone prefix, ConvBNReLU layers under `.net.0 / .net.1`,
    `pos_embed.pe` [1,400,512], `encoderA / encoderAB` in the score network, DataParallel prefix."""
    import re
    import torch
    out = {}
    for k, v in sd.items():
        k = k.replace("backbone.pe", "pos_embed.pe")
        k = re.sub(r"^backbone\.", "", k)
        k = re.sub(r"^(encodeAB?\.\d+)\.conv\.", r"\1.net.0.", k)
        k = re.sub(r"^(encodeAB?\.\d+)\.bn\.", r"\1.net.1.", k)
        if kind == 1:
            k = re.sub(r"^encode(AB?)\.", r"encoder\1.", k)
        out["module." + k] = v[None] if k == "pos_embed.pe" else v
        if k.endswith("running_var"):
            out["module." + k.replace("running_var", "num_batches_tracked")] = torch.tensor(7)
    return out


def test_weight_importer_upstream_names_and_exported_onnx(tmp_path):
    """(1) a checkpoint carrying the UPSTREAM parameter names imports to the same FPB2W file as the
    exporter; (2) an ONNX file shaped like a real torch.onnx export - BatchNorm folded into anonymous Conv constants,
    nn.Linear / packed MHA in-projection on 3-D activations turned into MatMul(x, W^T) + Add(bias) with anonymous
    transposed constants, one in-projection split into q / k / v products, no positional table - imports to it too.
    (`onnx` / `onnxscript` are absent from this image, so torch.onnx.export itself cannot run here: the graph is
    built with onnx_reader.write_onnx.)"""
    import numpy as np
    import torch
    from foundationpose_cpp_b200 import onnx_reader, weights
    for kind in (weights.KIND_REFINER, weights.KIND_SCORER):
        sd = weights.seeded_state_dict(kind, 21 + kind)
        ref = tmp_path / f"ref{kind}.fpb2w"
        (weights.export_refiner if kind == weights.KIND_REFINER else weights.export_scorer)(str(ref), sd)
        want = weights.read_weight_file(str(ref))
        # (1) upstream names
        up = _upstream_names(sd, kind)
        assert not any(k.startswith("module.backbone") for k in up) and "module.pos_embed.pe" in up
        assert any(".net.0.weight" in k for k in up)
        pth = tmp_path / f"up{kind}.pth"
        torch.save({"model": up}, pth)
        out = tmp_path / f"up{kind}.fpb2w"
        weights.import_weights(str(pth), str(out), kind)
        got = weights.read_weight_file(str(out))
        assert got.keys() == want.keys() and all(np.array_equal(got[k], want[k]) for k in want)
        # (2) export-shaped ONNX
        inits, nodes, n = {}, [], 0
        for name in weights._CONV_ORDER:
            bn = name.replace(".conv", ".bn") if name.endswith(".conv") else name.replace("conv", "bn")
            w, b = weights._fold(sd, name, bn)
            inits[f"onnx::Conv_{n}"], inits[f"onnx::Conv_{n + 1}"] = w.numpy(), b.numpy()
            nodes.append(("Conv", [f"x{n}", f"onnx::Conv_{n}", f"onnx::Conv_{n + 1}"], [f"y{n}"]))
            if "encodeA." in name:
                nodes.append(("Conv", [f"xb{n}", f"onnx::Conv_{n}", f"onnx::Conv_{n + 1}"], [f"yb{n}"]))
            n += 2
        for li, (name, o, i) in enumerate(weights._DENSE_ORDER[kind]):
            wk, bk = weights._dense_key(name, "weight"), weights._dense_key(name, "bias")
            w, b = sd[wk].numpy(), sd[bk].numpy()
            if o == 1536 and li == 0:  # one in-projection exported as three sliced products
                for j, part in enumerate("qkv"):
                    inits[f"onnx::MatMul_{n}"] = np.ascontiguousarray(w[512 * j:512 * (j + 1)].T)
                    inits[f"{name}.{part}_bias"] = b[512 * j:512 * (j + 1)].copy()
                    nodes.append(("MatMul", [f"t{n}", f"onnx::MatMul_{n}"], [f"m{n}"]))
                    nodes.append(("Add", [f"{name}.{part}_bias", f"m{n}"], [f"a{n}"]))
                    n += 1
            else:
                inits[f"onnx::MatMul_{n}"] = np.ascontiguousarray(w.T)
                inits[bk] = b
                nodes.append(("MatMul", [f"t{n}", f"onnx::MatMul_{n}"], [f"m{n}"]))
                nodes.append(("Add", [bk, f"m{n}"], [f"a{n}"]))
                n += 1
            nodes.append(("MatMul", [f"q{n}", f"k{n}"], [f"s{n}"]))  # activation x activation: must be ignored
        for k, v in sd.items():  # LayerNorm parameters keep their names in an export
            if ".norm" in k:
                inits[k] = v.numpy()
        onx = tmp_path / f"exp{kind}.onnx"
        onnx_reader.write_onnx(str(onx), inits, nodes)
        out = tmp_path / f"exp{kind}.fpb2w"
        weights.import_weights(str(onx), str(out), kind)
        got = weights.read_weight_file(str(out))
        assert got.keys() == want.keys()
        for k in want:
            np.testing.assert_allclose(got[k].astype(np.float32), want[k].astype(np.float32), rtol=2e-3, atol=1e-6)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU restatement timed as the reference arm): one JSON line with the
    contract's keys; no GPU needed."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    assert d["impl"] == "reference" and d["unit"] == "hyp/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Register pose-hypotheses/sec") and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    for k in ("n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d, k
