"""Host-side mirror of the reference's 6-DoF API over the C ABI (include/fpb200.h).

Names and argument meaning follow detection_6d_foundationpose/include/detection_6d_foundationpose/
foundationpose.hpp:16-105 and mesh_loader.hpp:15-93: `CreateFoundationPoseModel` returns an object
with `Register` / `Track` that return `(ok, pose)` — `ok` is the reference's bool, `pose` a 4x4
numpy array of the centred mesh in the OpenCV camera frame.  Like the reference, run-time failures
return False (message in `last_error`), construction failures raise.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import capi

CROP = 160
PACK_C = 8
STEM_PITCH = 164
STEM_PAD = 4


class BaseMeshLoader:
    """Mirror of detection_6d::BaseMeshLoader (mesh_loader.hpp:15-62)."""

    def GetName(self) -> str: raise NotImplementedError
    def GetMeshDiameter(self) -> float: raise NotImplementedError
    def GetMeshNumVertices(self) -> int: return len(self.GetMeshVertices())
    def GetMeshNumFaces(self) -> int: return len(self.GetMeshTriangleFaces())
    def GetMeshVertices(self) -> np.ndarray: raise NotImplementedError
    def GetMeshVertexNormals(self) -> np.ndarray: raise NotImplementedError
    def GetMeshTextureCoords(self) -> np.ndarray: raise NotImplementedError
    def GetMeshTriangleFaces(self) -> np.ndarray: raise NotImplementedError
    def GetMeshModelCenter(self) -> np.ndarray: raise NotImplementedError
    def GetOrientBounds(self) -> np.ndarray: raise NotImplementedError
    def GetObjectDimension(self) -> np.ndarray: raise NotImplementedError
    def GetTextureMap(self) -> np.ndarray: raise NotImplementedError


def compute_obb(vertices: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """ComputeOBB (assimp_mesh_loader.cpp:62-113): PCA of the vertex cloud.  Returns (orient_bounds 4x4, dimension 3):
    rotation = eigenvectors of the vertex covariance in ASCENDING eigenvalue order (Eigen's SelfAdjointEigenSolver
    convention, which numpy.linalg.eigh shares), translation = vertex mean, dimension = extent of the vertices along
    those axes.  The sign of each eigenvector is the solver's choice in the reference too (not pinned)."""
    v = np.asarray(vertices, np.float32)
    mean = v.sum(0, dtype=np.float32) / np.float32(len(v))
    diff = v - mean
    cov = (diff.T @ diff) / np.float32(len(v))
    _, rot = np.linalg.eigh(cov.astype(np.float32))
    rot = rot.astype(np.float32)
    tf = np.eye(4, dtype=np.float32)
    tf[:3, :3] = rot
    tf[:3, 3] = mean
    proj = v @ rot  # rows = rotation^T * vertex
    return tf, (proj.max(0) - proj.min(0)).astype(np.float32)


class ArrayMeshLoader(BaseMeshLoader):
    """Loader over in-memory arrays (e.g. synth.MeshData)."""

    def __init__(self, name, vertices, normals, texcoords, faces, texture, diameter=None, center=None):
        self._name = name
        self._v = np.ascontiguousarray(vertices, np.float32)
        self._n = np.ascontiguousarray(normals, np.float32)
        self._t = np.ascontiguousarray(texcoords, np.float32)
        self._f = np.ascontiguousarray(faces, np.int32)
        self._tex = np.ascontiguousarray(texture, np.uint8)
        if center is None:  # AABB midpoint, assimp_mesh_loader.cpp:179-180
            center = (self._v.min(0) + self._v.max(0)) / 2
        self._c = np.asarray(center, np.float32)
        if diameter is None:  # max pairwise vertex distance, assimp_mesh_loader.cpp:47-60 (on a subsample)
            pts = self._v[:: max(1, len(self._v) // 2000)]
            d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
            diameter = float(np.sqrt(d2.max()))
        self._d = float(diameter)
        self._obb, self._dim = compute_obb(self._v)

    @classmethod
    def from_mesh_data(cls, m):
        return cls(m.name, m.vertices, m.normals, m.texcoords, m.faces, m.texture, m.diameter, m.center)

    def GetName(self): return self._name
    def GetMeshDiameter(self): return self._d
    def GetMeshVertices(self): return self._v
    def GetMeshVertexNormals(self): return self._n
    def GetMeshTextureCoords(self): return self._t
    def GetMeshTriangleFaces(self): return self._f
    def GetMeshModelCenter(self): return self._c
    def GetOrientBounds(self): return self._obb
    def GetObjectDimension(self): return self._dim
    def GetTextureMap(self): return self._tex


def ConvertPoseMesh2BBox(pose_in_mesh: np.ndarray, loader: BaseMeshLoader) -> np.ndarray:
    """mesh_loader.hpp:75-81: T_bbox = T_mesh * translate(-center) * orient_bounds."""
    tf = np.eye(4, dtype=np.float32)
    tf[:3, 3] = -np.asarray(loader.GetMeshModelCenter(), np.float32)
    return np.asarray(pose_in_mesh, np.float32) @ tf @ np.asarray(loader.GetOrientBounds(), np.float32)


def mesh_diameter_gpu(vertices, device: int = 0) -> float:
    """CalcMeshDiameter (assimp_mesh_loader.cpp:47-60) on the device: maximum distance over all vertex pairs."""
    v = np.ascontiguousarray(vertices, np.float32)
    out = C.c_float(0)
    capi.check(capi.lib().fpb200_mesh_diameter(v.ctypes.data_as(C.c_void_p), int(v.shape[0]), int(device), C.byref(out)))
    return float(out.value)


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class FoundationPose:
    """Mirror of detection_6d::Base6DofDetectionModel backed by one fpb200 session."""

    def __init__(self, refiner_weights: str, scorer_weights: str, mesh_loaders: Sequence[BaseMeshLoader], K,
                 max_h: int = 1080, max_w: int = 1920, device: int = 0, max_hyps: int = 252, shard_rank: int = 0,
                 shard_world: int = 1, use_cuda_graph: bool = True):
        self._lib = capi.lib()
        self._keep = []
        meshes = (capi.Mesh * len(mesh_loaders))()
        for i, ld in enumerate(mesh_loaders):
            arrs = [np.ascontiguousarray(ld.GetMeshVertices(), np.float32),
                    np.ascontiguousarray(ld.GetMeshVertexNormals(), np.float32),
                    np.ascontiguousarray(ld.GetMeshTextureCoords(), np.float32),
                    np.ascontiguousarray(ld.GetMeshTriangleFaces(), np.int32),
                    np.ascontiguousarray(ld.GetTextureMap(), np.uint8)]
            if arrs[4].ndim != 3 or arrs[4].shape[2] != 3:
                raise ValueError("texture map must be [h,w,3] RGB")  # foundationpose_render.cpp:418-421
            name = ld.GetName().encode()
            self._keep += arrs + [name]
            m = meshes[i]
            m.name = name
            m.vertices, m.normals, m.texcoords, m.faces, m.texture = (_fptr(a) for a in arrs)
            m.num_vertices, m.num_faces = arrs[0].shape[0], arrs[3].shape[0]
            m.tex_h, m.tex_w = arrs[4].shape[0], arrs[4].shape[1]
            m.diameter = float(ld.GetMeshDiameter())
            c = np.asarray(ld.GetMeshModelCenter(), np.float32)
            m.center[0], m.center[1], m.center[2] = float(c[0]), float(c[1]), float(c[2])
        self._K = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
        opts = capi.Options()
        self._lib.fpb200_default_options(C.byref(opts))
        opts.device, opts.max_hyps, opts.shard_rank, opts.shard_world = device, max_hyps, shard_rank, shard_world
        opts.use_cuda_graph = int(bool(use_cuda_graph))
        self._h = C.c_void_p()
        capi.check(self._lib.fpb200_create(meshes, len(mesh_loaders), _fptr(self._K), max_h, max_w,
                                            refiner_weights.encode(), scorer_weights.encode(), C.byref(opts),
                                            C.byref(self._h)))
        self.last_error = ""
        self.shard_rank, self.shard_world = shard_rank, shard_world

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.fpb200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- reference API -------------------------------------------------------------------
    @staticmethod
    def _check_images(rgb, depth, mask=None):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        depth = np.ascontiguousarray(depth, np.float32)
        if rgb.ndim != 3 or rgb.shape[2] != 3 or depth.shape != rgb.shape[:2]:
            raise ValueError("rgb must be [h,w,3] u8 and depth [h,w] f32 of the same size")
        if mask is not None:
            mask = np.ascontiguousarray(mask, np.uint8)
            if mask.shape != depth.shape:
                raise ValueError("mask must be [h,w] u8 of the image size")
        return rgb, depth, mask

    def _ok(self, code: int) -> bool:
        if code == capi.OK:
            return True
        self.last_error = self._lib.fpb200_last_error().decode(errors="replace")
        return False

    def Register(self, rgb, depth, mask, target_name: str, refine_itr: int = 1) -> Tuple[bool, Optional[np.ndarray]]:
        """Base6DofDetectionModel::Register (foundationpose.hpp:36-41)."""
        try:
            rgb, depth, mask = self._check_images(rgb, depth, mask)
        except ValueError as e:  # the reference's CheckInputArguments returns false (foundationpose.cpp:155-179)
            self.last_error = str(e)
            return False, None
        out = np.zeros(16, np.float32)
        h, w = depth.shape
        # mask=None: reuse the mask already resident on the device (fails with "mask must not be null" if none is)
        ok = self._ok(self._lib.fpb200_register(self._h, _fptr(rgb), _fptr(depth), _fptr(mask) if mask is not None else None, h, w,
                                                 target_name.encode(), int(refine_itr), _fptr(out)))
        return ok, (out.reshape(4, 4).T.copy() if ok else None)

    def Track(self, rgb, depth, hyp_pose_in_mesh, target_name: str, refine_itr: int = 1):
        """Base6DofDetectionModel::Track (foundationpose.hpp:62-67)."""
        try:
            rgb, depth, _ = self._check_images(rgb, depth)
        except ValueError as e:
            self.last_error = str(e)
            return False, None
        hyp = np.ascontiguousarray(np.asarray(hyp_pose_in_mesh, np.float32).T)  # column-major
        out = np.zeros(16, np.float32)
        h, w = depth.shape
        ok = self._ok(self._lib.fpb200_track(self._h, _fptr(rgb), _fptr(depth), h, w, _fptr(hyp),
                                              target_name.encode(), int(refine_itr), _fptr(out)))
        return ok, (out.reshape(4, 4).T.copy() if ok else None)

    def track_async(self, rgb, depth, hyp_pose_in_mesh, target_name: str, refine_itr: int = 1):
        """Enqueue one Track without waiting (fpb200_track_async); rgb / depth must be C-contiguous arrays
        of the right dtype that stay alive until track_wait."""
        hyp = np.ascontiguousarray(np.asarray(hyp_pose_in_mesh, np.float32).T)
        h, w = depth.shape
        capi.check(self._lib.fpb200_track_async(self._h, _fptr(rgb), _fptr(depth), h, w, _fptr(hyp),
                                                target_name.encode(), int(refine_itr)))

    def track_wait(self):
        out = np.zeros(16, np.float32)
        capi.check(self._lib.fpb200_track_wait(self._h, _fptr(out)))
        return out.reshape(4, 4).T.copy()

    # ---- streamed tracking (fpb200_track_stream_*) ------------------------------------------
    def track_stream_begin(self, hyp_pose_in_mesh, target_name: str, refine_itr: int = 1):
        hyp = np.ascontiguousarray(np.asarray(hyp_pose_in_mesh, np.float32).T)
        capi.check(self._lib.fpb200_track_stream_begin(self._h, _fptr(hyp), target_name.encode(), int(refine_itr)))

    def track_stream_push(self, rgb, depth):
        """rgb / depth: C-contiguous u8 [h,w,3] / f32 [h,w] arrays that stay alive until the matching pop."""
        h, w = depth.shape
        capi.check(self._lib.fpb200_track_stream_push(self._h, _fptr(rgb), _fptr(depth), h, w))

    def track_stream_pop(self) -> np.ndarray:
        out = np.zeros(16, np.float32)
        capi.check(self._lib.fpb200_track_stream_pop(self._h, _fptr(out)))
        return out.reshape(4, 4).T.copy()

    # ---- stage API (tests, multi-GPU) ----------------------------------------------------
    def upload(self, rgb, depth):
        rgb, depth, _ = self._check_images(rgb, depth)
        capi.check(self._lib.fpb200_upload(self._h, _fptr(rgb), _fptr(depth), depth.shape[0], depth.shape[1]))

    def sample_hypotheses(self, mask, target_name: str):
        """mask=None reuses the mask already resident on the device."""
        if mask is None:
            capi.check(self._lib.fpb200_sample_hypotheses(self._h, None, target_name.encode()))
            return
        mask = np.ascontiguousarray(mask, np.uint8)
        capi.check(self._lib.fpb200_sample_hypotheses(self._h, _fptr(mask), target_name.encode()))

    def set_hypotheses(self, poses, target_name: str):
        p = np.ascontiguousarray(np.asarray(poses, np.float32).transpose(0, 2, 1))
        capi.check(self._lib.fpb200_set_hypotheses(self._h, _fptr(p), p.shape[0], target_name.encode()))

    def get_hypotheses(self) -> np.ndarray:
        buf = np.zeros((4096, 16), np.float32)
        n = C.c_int(0)
        capi.check(self._lib.fpb200_get_hypotheses(self._h, _fptr(buf), 4096, C.byref(n)))
        return buf[: n.value].reshape(-1, 4, 4).transpose(0, 2, 1).copy()

    def render_and_transform(self, crop_ratio: float):
        capi.check(self._lib.fpb200_render_and_transform(self._h, C.c_float(crop_ratio)))

    def refine_step(self):
        capi.check(self._lib.fpb200_refine_step(self._h))

    def score_features(self):
        capi.check(self._lib.fpb200_score_features(self._h))

    def feature_buffers(self):
        lp, ap = C.c_void_p(), C.c_void_p()
        nl, off, na = C.c_int(), C.c_int(), C.c_int()
        capi.check(self._lib.fpb200_feature_buffers(self._h, C.byref(lp), C.byref(nl), C.byref(off), C.byref(ap),
                                                     C.byref(na)))
        return lp.value, nl.value, off.value, ap.value, na.value

    def score_finish(self):
        idx = C.c_int(-1)
        out = np.zeros(16, np.float32)
        capi.check(self._lib.fpb200_score_finish(self._h, C.byref(idx), _fptr(out)))
        return idx.value, out.reshape(4, 4).T.copy()

    def read_buffer(self, name: str, dtype, count: int) -> np.ndarray:
        buf = np.zeros(count, dtype)
        written = C.c_size_t(0)
        capi.check(self._lib.fpb200_read_buffer(self._h, name.encode(), _fptr(buf), C.c_size_t(buf.nbytes),
                                                 C.byref(written)))
        return buf[: written.value // buf.itemsize]

    def read_packed(self, which: str, n: int) -> np.ndarray:
        """'render_input' / 'transf_input' as f32 [n,160,160,6] (pad pixels / channels stripped)."""
        raw = self.read_buffer(which, np.float16, n * CROP * STEM_PITCH * PACK_C)
        return raw.reshape(n, CROP, STEM_PITCH, PACK_C)[:, :, STEM_PAD:, :6].astype(np.float32)

    def debug_options(self, rast_ids: bool = False, render_split: int = -1):
        """Test switches (fpb200_debug_options): write the visibility id map, force a render path."""
        capi.check(self._lib.fpb200_debug_options(self._h, int(bool(rast_ids)), int(render_split)))

    def time_stage(self, stage: int, repeat: int = 20) -> float:
        """Average device ms of one stage (0 render, 1 warp, 2 exchange) on the current hypotheses."""
        ms = C.c_float(0)
        capi.check(self._lib.fpb200_debug_time_stage(self._h, int(stage), int(repeat), C.byref(ms)))
        return ms.value

    def last_timings(self) -> List[float]:
        ms = (C.c_float * 8)()
        n = self._lib.fpb200_last_timings(self._h, ms, 8)
        return [ms[i] for i in range(max(n, 0))]

    def flops_per_hypothesis(self) -> Tuple[float, float]:
        a, b = C.c_double(), C.c_double()
        capi.check(self._lib.fpb200_flops_per_hypothesis(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def last_launch_count(self) -> int:
        return int(self._lib.fpb200_last_launch_count(self._h))

    def launch_count_total(self) -> int:
        self._lib.fpb200_launch_count_total.restype = C.c_longlong
        return int(self._lib.fpb200_launch_count_total(self._h))

    # ---- asynchronous use ----------------------------------------------------------------
    def stream_ptr(self) -> int:
        self._lib.fpb200_stream.restype = C.c_void_p
        return int(self._lib.fpb200_stream(self._h) or 0)

    def set_async(self, on: bool):
        capi.check(self._lib.fpb200_set_async(self._h, int(bool(on))))

    def register_async(self, target_name: str, refine_itr: int = 1):
        """Whole Register on the resident frame, enqueued without synchronising (see fpb200_register_async)."""
        capi.check(self._lib.fpb200_register_async(self._h, target_name.encode(), int(refine_itr)))

    def score_finish_async(self):
        capi.check(self._lib.fpb200_score_finish_async(self._h))

    def wait(self):
        idx = C.c_int(-1)
        out = np.zeros(16, np.float32)
        capi.check(self._lib.fpb200_wait(self._h, C.byref(idx), _fptr(out)))
        return idx.value, out.reshape(4, 4).T.copy()

    # ---- multi-GPU exchange over peer memory (fpb200_exchange_*) -----------------------------
    def exchange_export(self) -> bytes:
        """CUDA IPC handle (64 bytes) of this rank's exchange region."""
        buf = (C.c_ubyte * capi.IPC_HANDLE_BYTES)()
        capi.check(self._lib.fpb200_exchange_export(self._h, buf))
        return bytes(buf)

    def exchange_connect(self, handles: Sequence[bytes]):
        """handles: the exported handle of every rank, in rank order (this rank's own entry is ignored)."""
        if len(handles) != self.shard_world or any(len(h) != capi.IPC_HANDLE_BYTES for h in handles):
            raise ValueError(f"need {self.shard_world} handles of {capi.IPC_HANDLE_BYTES} bytes")
        blob = b"".join(handles)
        capi.check(self._lib.fpb200_exchange_connect(self._h, (C.c_ubyte * len(blob)).from_buffer_copy(blob)))

    def set_gemm_profile(self, on: bool):
        capi.check(self._lib.fpb200_set_gemm_profile(self._h, int(bool(on))))

    def gemm_profile(self):
        ms, fl, n = C.c_double(), C.c_double(), C.c_longlong()
        capi.check(self._lib.fpb200_gemm_profile(self._h, C.byref(ms), C.byref(fl), C.byref(n)))
        return ms.value, fl.value, n.value


def connect_local(models: Sequence[FoundationPose]):
    """Wire the exchange of `world` sharded sessions that live in this process (rank order)."""
    arr = (C.c_void_p * len(models))(*[m._h for m in models])
    capi.check(capi.lib().fpb200_exchange_connect_local(arr, len(models)))


def CreateFoundationPoseModel(refiner_weights: str, scorer_weights: str, mesh_loaders, K, max_h: int = 1080,
                              max_w: int = 1920, **kw) -> FoundationPose:
    """CreateFoundationPoseModel (foundationpose.hpp:99-105); the weight files stand where the
    reference passes two TensorRT BaseInferCore objects."""
    return FoundationPose(refiner_weights, scorer_weights, mesh_loaders, K, max_h, max_w, **kw)
