"""Seeded synthetic meshes and RGB-D scenes (no dataset is reachable offline).

The reference's demo data (mustard0, test_data/download.md) and its assimp loader are not
available; these generators produce the `BaseMeshLoader` fields
(detection_6d_foundationpose/include/detection_6d_foundationpose/mesh_loader.hpp:15-62) and frames
in the conventions of simple_tests/include/tests/help_func.hpp:14-32 (rgb u8 RGB, depth f32 metres,
mask u8).  The scene is ray-cast analytically (ellipsoid + background plane), independently of any
rasteriser in this repository.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np


@dataclass
class MeshData:
    name: str
    vertices: np.ndarray   # [V,3] f32 (un-centred, as a loader returns them)
    normals: np.ndarray    # [V,3] f32
    texcoords: np.ndarray  # [V,2] f32
    faces: np.ndarray      # [T,3] i32
    texture: np.ndarray    # [th,tw,3] u8 RGB
    diameter: float
    center: np.ndarray     # [3] f32 (AABB midpoint)
    axes: np.ndarray       # ellipsoid semi-axes used by the scene generator


def _smooth_texture(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    base = rng.random((size // 16, size // 16, 3))
    tex = np.kron(base, np.ones((16, 16, 1)))
    yy, xx = np.mgrid[0:size, 0:size]
    stripes = 0.5 + 0.5 * np.sin(xx / size * 2 * math.pi * 6)[..., None] * np.cos(yy / size * 2 * math.pi * 4)[..., None]
    tex = 0.65 * tex + 0.35 * stripes
    tex += rng.normal(0, 0.03, tex.shape)
    return np.clip(tex * 255, 0, 255).astype(np.uint8)


def uv_ellipsoid(name: str = "object", slices: int = 100, stacks: int = 51, axes=(0.05, 0.05, 0.09),
                 tex_size: int = 512, seed: int = 0, offset=(0.01, -0.02, 0.03)) -> MeshData:
    """UV ellipsoid with `2*slices*(stacks-1)` triangles (100 x 51 -> exactly 10 000); the seam
    column is duplicated so that every vertex carries one UV, as BaseMeshLoader requires."""
    ax = np.asarray(axes, np.float64)
    verts, uvs = [], []
    # rows 1..stacks-1 are rings of slices+1 vertices (seam duplicated); poles are single vertices
    for j in range(1, stacks):
        th = math.pi * j / stacks
        for i in range(slices + 1):
            ph = 2 * math.pi * i / slices
            verts.append((math.sin(th) * math.cos(ph), math.sin(th) * math.sin(ph), math.cos(th)))
            uvs.append((i / slices, j / stacks))
    north, south = len(verts), len(verts) + 1
    verts += [(0, 0, 1), (0, 0, -1)]
    uvs += [(0.5, 0.0), (0.5, 1.0)]
    unit = np.asarray(verts, np.float64)
    faces = []
    ring = slices + 1
    for i in range(slices):
        faces.append((north, i, i + 1))
    for j in range(stacks - 2):
        for i in range(slices):
            a, b = j * ring + i, j * ring + i + 1
            c, d = (j + 1) * ring + i, (j + 1) * ring + i + 1
            faces.append((a, c, b))
            faces.append((b, c, d))
    last = (stacks - 2) * ring
    for i in range(slices):
        faces.append((south, last + i + 1, last + i))
    pos = unit * ax
    nrm = unit / ax
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    off = np.asarray(offset, np.float64)
    pos = pos + off  # loaders return un-centred vertices
    lo, hi = pos.min(0), pos.max(0)
    center = ((lo + hi) / 2).astype(np.float32)
    diameter = float(2 * ax.max())  # max pairwise distance of an ellipsoid = longest axis
    return MeshData(name, pos.astype(np.float32), nrm.astype(np.float32), np.asarray(uvs, np.float32),
                    np.asarray(faces, np.int32), _smooth_texture(tex_size, seed), diameter, center,
                    ax.astype(np.float32))


def random_rotation(seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)


DEFAULT_K = np.array([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]], np.float32)


def make_scene(mesh: MeshData, K=DEFAULT_K, h: int = 480, w: int = 640, t=(0.02, -0.01, 0.60), rot_seed: int = 1,
               noise_seed: int = 1, background_z: float = 1.2):
    """Ray-cast the textured ellipsoid at pose (R,t) of the CENTRED mesh over a textured background
    plane.  Returns rgb u8 [h,w,3], depth f32 [h,w] (1 mm noise, 2 % dropouts), mask u8, pose 4x4."""
    K = np.asarray(K, np.float64)
    R = random_rotation(rot_seed)
    t = np.asarray(t, np.float64)
    ax = mesh.axes.astype(np.float64)
    # the centred mesh is the ellipsoid centred at (offset - center) in its own frame
    c_obj = (mesh.vertices.astype(np.float64).mean(0) * 0)  # ellipsoid centre before centring == offset
    ell_center = (mesh.vertices.max(0).astype(np.float64) + mesh.vertices.min(0).astype(np.float64)) / 2 - mesh.center
    us, vs = np.meshgrid(np.arange(w), np.arange(h))
    rays = np.stack([(us - K[0, 2]) / K[0, 0], (vs - K[1, 2]) / K[1, 1], np.ones_like(us, dtype=np.float64)], -1)
    # transform rays into the object frame: p_obj = R^T (p_cam - t)
    o = (R.T @ (-t)) - ell_center
    d = rays @ R  # (R^T d) for row vectors
    oa, da = o / ax, d / ax
    A = np.sum(da * da, -1)
    B = 2 * np.sum(da * oa, -1)
    C = np.sum(oa * oa) - 1
    disc = B * B - 4 * A * C
    hit = disc > 0
    lam = np.where(hit, (-B - np.sqrt(np.where(hit, disc, 0))) / (2 * A), np.inf)
    hit &= lam > 0
    p_obj = o + d * lam[..., None]
    unit = np.where(hit[..., None], p_obj / ax, 0)
    unit = np.clip(unit, -1, 1)
    uu = (np.arctan2(unit[..., 1], unit[..., 0]) / (2 * math.pi)) % 1.0
    vv = np.arccos(unit[..., 2]) / math.pi
    th, tw, _ = mesh.texture.shape
    # (u, 1-v) flip + texel lookup consistent with the renderer's convention
    tx = np.clip((uu * tw).astype(int), 0, tw - 1)
    ty = np.clip(((1 - vv) * th).astype(int), 0, th - 1)
    tex = mesh.texture[ty, tx].astype(np.float64) / 255.0
    n_obj = unit / ax
    n_obj /= np.maximum(np.linalg.norm(n_obj, axis=-1, keepdims=True), 1e-9)
    n_cam = n_obj @ R.T
    diffuse = np.clip(-n_cam[..., 2], 0, 1)
    col_obj = np.clip(tex * (0.8 + 0.5 * diffuse[..., None]), 0, 1)
    rng = np.random.default_rng(noise_seed)
    bg = (np.kron(rng.random((-(-h // 16), -(-w // 16), 3)), np.ones((16, 16, 1))) * 0.6 + 0.2)[:h, :w]
    rgb = np.where(hit[..., None], col_obj, bg)
    depth = np.where(hit, lam, background_z).astype(np.float64)  # rays have z = 1 -> lam is depth
    depth += rng.normal(0, 0.001, depth.shape)
    drop = rng.random(depth.shape) < 0.02
    depth[drop] = 0
    pose = np.eye(4, dtype=np.float32)
    pose[:3, :3] = R
    pose[:3, 3] = t
    del c_obj
    return ((rgb * 255).astype(np.uint8), depth.astype(np.float32), (hit * 255).astype(np.uint8), pose)


def perturb_pose(pose: np.ndarray, deg: float, trans: float, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    a = math.radians(deg)
    kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    dr = np.eye(3) + math.sin(a) * kx + (1 - math.cos(a)) * (kx @ kx)
    out = pose.astype(np.float64).copy()
    out[:3, :3] = dr @ out[:3, :3]
    dv = rng.normal(size=3)
    out[:3, 3] += dv / np.linalg.norm(dv) * trans
    return out.astype(np.float32)
