"""Shared fixtures helpers: seeded weights, synthetic scene, oracle-side mesh dictionary."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from foundationpose_cpp_b200 import synth, weights  # noqa: E402
from oracle import nets  # noqa: E402

F = np.float32


def oracle_mesh(m: synth.MeshData) -> dict:
    """What FoundationPoseRenderer::LoadTexturedMesh uploads (foundationpose_render.cpp:381-509)."""
    tc = m.texcoords.astype(F).copy()
    tc[:, 1] = F(1) - tc[:, 1]
    return {
        "vertices": (m.vertices.astype(F) - m.center.astype(F)).astype(F),
        "normals": m.normals.astype(F),
        "texcoords": tc,
        "faces": m.faces.astype(np.int32),
        "texture": (m.texture.astype(F) * F(1.0 / 255.0)).astype(F),
        "diameter": float(m.diameter),
    }


_NETS = {}


def seeded_nets():
    if not _NETS:
        torch.manual_seed(0)
        _NETS["refiner"] = nets.seeded_init(nets.RefineNet(), 2)
        _NETS["scorer"] = nets.seeded_init(nets.ScoreNet(), 3)
    return _NETS["refiner"], _NETS["scorer"]


def write_seeded_weights(dirpath: str):
    r, s = seeded_nets()
    rp, sp = os.path.join(dirpath, "refiner.fpb2w"), os.path.join(dirpath, "scorer.fpb2w")
    if not os.path.exists(rp):
        weights.export_refiner(rp, r.state_dict())
    if not os.path.exists(sp):
        weights.export_scorer(sp, s.state_dict())
    return rp, sp
