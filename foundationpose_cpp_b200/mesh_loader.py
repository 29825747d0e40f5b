"""OBJ / MTL mesh loader producing the `BaseMeshLoader` fields without assimp.

Mirrors what the reference's AssimpMeshLoader computes
(detection_6d_foundationpose/src/mesh_loader/assimp_mesh_loader.cpp:159-228): triangulated faces of the
first mesh with identical (position, texcoord, normal) corners joined into one vertex (aiProcess_Triangulate |
JoinIdenticalVertices, :168-170), diameter = maximum pair-wise vertex distance (:47-60), centre = AABB
midpoint (:179-180), texture = the material's diffuse map converted to RGB, or a 2x2 grey (100,100,100)
image when there is none (:207-223).  Texture coordinates are required, as in the reference (:182-185).
"""
from __future__ import annotations

import os
from typing import Dict, List, Tuple

import numpy as np

from .api import ArrayMeshLoader


def _parse_mtl(path: str) -> Dict[str, str]:
    """material name -> diffuse texture path (map_Kd)."""
    out, cur = {}, None
    if not os.path.exists(path):
        return out
    base = os.path.dirname(path)
    with open(path, "r", errors="replace") as f:
        for line in f:
            t = line.split()
            if not t:
                continue
            if t[0] == "newmtl":
                cur = " ".join(t[1:])
            elif t[0] == "map_Kd" and cur is not None:
                out[cur] = os.path.join(base, t[-1])
    return out


def max_pairwise_distance(v: np.ndarray, chunk: int = 2048) -> float:
    """Exact O(V^2) diameter like the reference, evaluated in chunks."""
    v = np.asarray(v, np.float64)
    best = 0.0
    sq = (v * v).sum(1)
    for i in range(0, len(v), chunk):
        d2 = sq[i:i + chunk, None] + sq[None, :] - 2.0 * (v[i:i + chunk] @ v.T)
        best = max(best, float(d2.max()))
    return float(np.sqrt(max(best, 0.0)))


def load_obj(path: str) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Returns vertices [V,3], normals [V,3], texcoords [V,2], faces [T,3] i32, texture [h,w,3] u8 RGB."""
    pos: List[Tuple[float, ...]] = []
    tex: List[Tuple[float, ...]] = []
    nrm: List[Tuple[float, ...]] = []
    corners: Dict[Tuple[int, int, int], int] = {}
    out_v, out_t, out_n, faces = [], [], [], []
    mtllib, usemtl = None, None
    with open(path, "r", errors="replace") as f:
        for line in f:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "v":
                pos.append(tuple(float(x) for x in t[1:4]))
            elif t[0] == "vt":
                tex.append((float(t[1]), float(t[2]) if len(t) > 2 else 0.0))
            elif t[0] == "vn":
                nrm.append(tuple(float(x) for x in t[1:4]))
            elif t[0] == "mtllib":
                mtllib = " ".join(t[1:])
            elif t[0] == "usemtl" and usemtl is None:
                usemtl = " ".join(t[1:])
            elif t[0] == "f":
                idx = []
                for c in t[1:]:
                    p = (c.split("/") + ["", ""])[:3]
                    vi = int(p[0])
                    ti = int(p[1]) if p[1] else 0
                    ni = int(p[2]) if p[2] else 0
                    vi = vi - 1 if vi > 0 else len(pos) + vi
                    ti = ti - 1 if ti > 0 else (len(tex) + ti if ti < 0 else -1)
                    ni = ni - 1 if ni > 0 else (len(nrm) + ni if ni < 0 else -1)
                    if ti < 0:
                        raise ValueError(f"{path}: faces without texture coordinates are not supported "
                                         "(the reference requires HasTextureCoords)")
                    key = (vi, ti, ni)
                    if key not in corners:
                        corners[key] = len(out_v)
                        out_v.append(pos[vi])
                        out_t.append(tex[ti])
                        out_n.append(nrm[ni] if ni >= 0 else (0.0, 0.0, 0.0))
                    idx.append(corners[key])
                for k in range(1, len(idx) - 1):  # fan triangulation
                    faces.append((idx[0], idx[k], idx[k + 1]))
    if not faces:
        raise ValueError(f"{path}: no faces")
    v = np.asarray(out_v, np.float32)
    n = np.asarray(out_n, np.float32)
    fc = np.asarray(faces, np.int32)
    if not np.any(n):  # no normals in the file: area-weighted vertex normals
        fn = np.cross(v[fc[:, 1]] - v[fc[:, 0]], v[fc[:, 2]] - v[fc[:, 0]])
        n = np.zeros_like(v)
        for k in range(3):
            np.add.at(n, fc[:, k], fn)
    ln = np.linalg.norm(n, axis=1, keepdims=True)
    n = (n / np.where(ln > 0, ln, 1)).astype(np.float32)
    texture = np.full((2, 2, 3), 100, np.uint8)
    if mtllib:
        maps = _parse_mtl(os.path.join(os.path.dirname(path), mtllib))
        tp = maps.get(usemtl) if usemtl in maps else (next(iter(maps.values())) if maps else None)
        if tp and os.path.exists(tp):
            import cv2
            img = cv2.imread(tp, cv2.IMREAD_COLOR)
            if img is not None:
                texture = np.ascontiguousarray(img[..., ::-1])  # BGR -> RGB
    return v, n, np.asarray(out_t, np.float32), fc, texture


class ObjMeshLoader(ArrayMeshLoader):
    def __init__(self, name: str, path: str):
        v, n, t, f, tex = load_obj(path)
        super().__init__(name, v, n, t, f, tex, diameter=max_pairwise_distance(v),
                         center=(v.min(0) + v.max(0)) / 2)


def CreateAssimpMeshLoader(name: str, mesh_file_path: str) -> ObjMeshLoader:
    """Name kept from mesh_loader.hpp:92-93; only Wavefront OBJ (+MTL, image via cv2) is supported."""
    return ObjMeshLoader(name, mesh_file_path)


def save_obj(path: str, vertices, normals, texcoords, faces, texture=None) -> None:
    """Writes OBJ (+ MTL + PNG when a texture is given); used by tests and tools."""
    base = os.path.splitext(path)[0]
    with open(path, "w") as f:
        if texture is not None:
            f.write(f"mtllib {os.path.basename(base)}.mtl\nusemtl material_0\n")
        for p in vertices:
            f.write(f"v {p[0]:.9g} {p[1]:.9g} {p[2]:.9g}\n")
        for t in texcoords:
            f.write(f"vt {t[0]:.9g} {t[1]:.9g}\n")
        for n in normals:
            f.write(f"vn {n[0]:.9g} {n[1]:.9g} {n[2]:.9g}\n")
        for a, b, c in np.asarray(faces) + 1:
            f.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
    if texture is not None:
        import cv2
        cv2.imwrite(base + ".png", np.asarray(texture)[..., ::-1])
        with open(base + ".mtl", "w") as f:
            f.write(f"newmtl material_0\nKd 1 1 1\nmap_Kd {os.path.basename(base)}.png\n")
