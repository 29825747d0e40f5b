"""CPU known-answer tests of the oracle (oracle/).  The reference's own tests hold no golden vectors
for this path (simple_tests/src/test_foundationpose.cpp asserts only booleans), so these are
hand-computed answers for the formulas cited in oracle/fp_oracle.py, plus an independent
cross-check of the crop/warp against cv2.warpPerspective."""
import math

import numpy as np
import pytest

from oracle import fp_oracle as O

F = np.float32
K = np.array([[600.0, 0, 320.0], [0, 600.0, 240.0], [0, 0, 1.0]], F)


def test_rotation_grid_structure():
    g = O.rotation_grid()
    assert g.shape == (252, 4, 4)  # 42 icosphere vertices x 6 in-plane steps
    assert O.icosphere_vertices().shape == (42, 3)
    r = g[:, :3, :3].astype(np.float64)
    np.testing.assert_allclose(r @ r.transpose(0, 2, 1), np.broadcast_to(np.eye(3), (252, 3, 3)), atol=2e-6)
    np.testing.assert_allclose(np.linalg.det(r), 1.0, atol=1e-5)
    # camera sits on the unit sphere looking at the origin: ob_in_cam translation = (0, 0, 1)
    np.testing.assert_allclose(g[:, :3, 3], np.broadcast_to([0, 0, 1], (252, 3)), atol=2e-6)
    # in-plane minor order: consecutive entries of one view differ by a 60 degree rotation about z (camera)
    rel = r[1] @ r[0].T
    ang = math.degrees(math.acos((np.trace(rel) - 1) / 2))
    assert abs(ang - 60) < 1e-3 and abs(abs(rel[2, 2]) - 1) < 1e-5
    # first icosphere vertex is (-1, phi, 0)/norm
    phi = (1 + 5 ** 0.5) / 2
    np.testing.assert_allclose(O.icosphere_vertices()[0], np.array([-1, phi, 0]) / np.hypot(1, phi), atol=1e-6)


def test_guess_translation_known_answer():
    depth = np.zeros((480, 640), F)
    mask = np.zeros((480, 640), np.uint8)
    mask[100:201, 300:341] = 255           # u in [300,340], v in [100,200] -> centre (320, 150)
    depth[100:201, 300:341] = 0.5
    depth[100, 300:341] = 0.7              # 41 outliers do not move the median
    depth[150, 320] = 0.0                  # invalid pixel is ignored
    c = O.guess_translation(depth, mask, K)
    np.testing.assert_allclose(c, [(320 - 320) / 600 * 0.5, (150 - 240) / 600 * 0.5, 0.5], atol=1e-7)
    assert O.guess_translation(depth, np.zeros_like(mask), K) is None
    assert O.guess_translation(np.zeros_like(depth), mask, K) is None
    # even count: mean of the two middle values
    d2 = np.zeros((4, 4), F); m2 = np.zeros((4, 4), np.uint8)
    d2[0, :4] = [0.2, 0.4, 0.6, 0.8]; m2[0, :4] = 1
    assert abs(O.guess_translation(d2, m2, K)[2] - 0.5) < 1e-7


def test_depth_filters_small_cases():
    d = np.full((9, 9), 0.5, F)
    d[4, 4] = 0.05                         # below 0.1 -> removed
    d[0, 0] = 150.0                        # beyond zfar -> removed
    e = O.erode_depth(d)
    assert e[4, 4] == 0 and e[0, 0] == 0 and e[2, 2] == F(0.5)
    spike = np.full((9, 9), 0.5, F); spike[4, 4] = 0.6  # isolated pixel: > 80 % of neighbours differ by > 1 mm
    assert O.erode_depth(spike)[4, 4] == 0
    b = O.bilateral_filter_depth(np.full((9, 9), 0.5, F))
    np.testing.assert_allclose(b, 0.5, rtol=1e-6)
    assert np.all(O.bilateral_filter_depth(np.zeros((5, 5), F)) == 0)


def test_crop_window_hand_computed():
    pose = np.eye(4, dtype=F)
    pose[:3, 3] = [0.02, -0.01, 0.6]
    tf, box = O.crop_window_tf(pose[None], K, 1.2, 0.2)
    # centre (340, 230); radius = 600 * 0.12 / 0.6 = 120 -> window [220,460] x [110,350], scale 160/240
    np.testing.assert_allclose(box[0], [220, 460, 110, 350])
    np.testing.assert_allclose(tf[0], [[2 / 3, 0, -220 * 2 / 3], [0, 2 / 3, -110 * 2 / 3], [0, 0, 1]], rtol=1e-6)
    bb = O.bbox2d_from_tf(tf)
    np.testing.assert_allclose(bb[0], [220, 110, 220 + 159 * 1.5, 110 + 159 * 1.5], rtol=1e-6)
    # half-pixel rounding is std::round (away from zero), not banker's rounding
    K2 = K.copy(); K2[0, 2] = 320.5         # u0 = 320.5 exactly; r = 0.125, z = 0.5 -> radius 150 exactly
    pose[:3, 3] = [0.0, 0.0, 0.5]
    _, box = O.crop_window_tf(pose[None], K2, 1.0, 0.25)
    assert box[0][0] == 171 and box[0][1] == 471   # 170.5 -> 171, 470.5 -> 471 (rint would give 170, 470)


def test_projection_matrix_values():
    p = O.projection_matrix(K, 480, 640)
    np.testing.assert_allclose(p[0], [2 * 600 / 640, 0, (640 - 2 * 320) / 640, 0], atol=1e-7)
    np.testing.assert_allclose(p[1], [0, 2 * 600 / 480, (2 * 240 - 480) / 480, 0], atol=1e-7)
    np.testing.assert_allclose(p[2], [0, 0, -(100.1) / 99.9, -2 * 10 / 99.9], rtol=1e-6)
    np.testing.assert_allclose(p[3], [0, 0, -1, 0])


def test_raster_two_triangles_fill_rule_and_depth():
    # clip-space quad covering NDC [-0.5,0.5]^2 at w=1, split along the diagonal; a nearer small triangle on top
    v = np.array([[-0.5, -0.5, 0.2, 1], [0.5, -0.5, 0.2, 1], [0.5, 0.5, 0.2, 1], [-0.5, 0.5, 0.2, 1],
                  [-0.25, -0.25, -0.5, 1], [0.25, -0.25, -0.5, 1], [0.0, 0.25, -0.5, 1]], F)
    f = np.array([[0, 1, 2], [0, 2, 3], [4, 5, 6]], np.int32)
    ids, nclip = O.rasterize_ids(v, f, 16, 16)
    assert nclip == 0
    # pixel centres are at (px + 0.5) / 8 - 1; the quad spans pixels 4..11 in both axes
    covered = ids > 0
    assert covered.sum() >= 64 and covered[4:12, 4:12].all() and not covered[:4].any() and not covered[:, 12:].any()
    # every covered pixel belongs to exactly one triangle of the quad (shared diagonal: no holes, no overlaps)
    assert set(np.unique(ids[4:12, 4:12])) <= {1, 2, 3}
    assert (ids == 3).sum() > 0 and ids[8, 8] == 3  # nearer triangle wins at the centre
    # equal depth: the later triangle wins
    f2 = np.array([[0, 1, 2], [0, 1, 2]], np.int32)
    ids2, _ = O.rasterize_ids(v, f2, 16, 16)
    assert set(np.unique(ids2)) == {0, 2}
    # degenerate and fully outside triangles draw nothing
    v3 = np.array([[0, 0, 0, 1], [0, 0, 0, 1], [0.1, 0.1, 0, 1], [2, 2, 0, 1], [3, 2, 0, 1], [2, 3, 0, 1]], F)
    ids3, _ = O.rasterize_ids(v3, np.array([[0, 1, 2], [3, 4, 5]], np.int32), 16, 16)
    assert ids3.sum() == 0
    # a triangle crossing the near plane takes cudaraster's polygon clipper: the part with z/w >= -1 is
    # drawn.  Vertices (0,0,0), (0.5,0,-2), (0,0.5,0): z = -4x, so the kept part is x <= 0.25.
    v4 = np.array([[0, 0, 0, 1], [0.5, 0, -2, 1], [0, 0.5, 0, 1]], F)
    ids4, nclip4 = O.rasterize_ids(v4, np.array([[0, 1, 2]], np.int32), 16, 16)
    assert nclip4 == 1
    ys, xs = np.nonzero(ids4)
    cx = (xs + 0.5) / 8 - 1  # NDC x of the covered pixel centres
    assert len(xs) > 0 and cx.max() <= 0.25 and cx.min() >= 0
    # the unclipped triangle (same footprint, z inside the range) covers strictly more
    v5 = v4.copy()
    v5[1, 2] = 0
    ids5, nclip5 = O.rasterize_ids(v5, np.array([[0, 1, 2]], np.int32), 16, 16)
    assert nclip5 == 0 and (ids5 > 0).sum() > (ids4 > 0).sum() and ((ids4 > 0) <= (ids5 > 0)).all()


def test_barycentrics_and_interpolate():
    v = np.array([[-1, -1, 0, 1], [1, -1, 0, 1], [-1, 1, 0, 1]], F)
    f = np.array([[0, 1, 2]], np.int32)
    ids, _ = O.rasterize_ids(v, f, 8, 8)
    rast = O.barycentrics(v, f, ids)
    attr = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1]], F)
    out = O.interpolate(attr, f, rast)
    m = ids > 0
    np.testing.assert_allclose(out[m].sum(-1), 1.0, atol=1e-6)
    # pixel (0,0): centre at NDC (-0.875,-0.875): b1 = (x+1)/2, b2 = (y+1)/2
    np.testing.assert_allclose(out[0, 0], [1 - 0.0625 * 2, 0.0625, 0.0625], atol=1e-6)
    assert np.all(out[~m] == 0)


def test_texture_wrap_and_bilinear():
    tex = np.zeros((2, 2, 3), F)
    tex[0, 0] = 1.0                       # texel (u=0.25, v=0.25) is white
    uv = np.array([[[0.25, 0.25], [0.5, 0.25], [0.0, 0.25], [1.25, 1.25]]], F)
    c = O.texture_linear(tex, uv)[0]
    np.testing.assert_allclose(c[0], 1.0)              # texel centre
    np.testing.assert_allclose(c[1], 0.5)              # halfway to the black neighbour
    np.testing.assert_allclose(c[2], 0.5)              # u = 0 wraps to the neighbour on the left (black) / white
    np.testing.assert_allclose(c[3], 1.0)              # integer part discarded (wrap)


def test_warp_matches_cv2_on_integer_and_ramp():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(0)
    rgb = rng.integers(0, 256, (480, 640, 3), dtype=np.uint8)
    xyz = rng.random((480, 640, 3)).astype(F)
    pose = np.eye(4, dtype=F); pose[:3, 3] = [0.02, -0.01, 0.6]
    tf, box = O.crop_window_tf(pose[None], K, 1.2, 0.2)
    got_rgb, got_xyz = O.warp_crop(rgb, xyz, tf[0], box[0])
    ref = cv2.warpPerspective(rgb, tf[0].astype(np.float64), (160, 160), flags=cv2.INTER_LINEAR,
                              borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    # cv2 quantises bilinear weights to 1/32: allow 255/32 ~ 8 levels worst case, 1 level typical
    diff = np.abs(got_rgb * 255 - ref.astype(F))
    assert np.median(diff) <= 1.0 and diff.max() <= 9.0
    # output pixels that map to integer source positions are exact copies
    assert np.array_equal((got_rgb[0:160:2, 0:160:2] * 255).round().astype(np.uint8), rgb[110:350:3, 220:460:3])
    assert np.array_equal(got_xyz[0:160:2, 0:160:2], xyz[110:350:3, 220:460:3])
    # nearest vs cv2 on a window whose scale (80/103) produces .5 ties only on two columns / rows:
    # everywhere else the two must agree exactly (the tie convention, floor(c + 0.5), is CV-CUDA's)
    pose[:3, 3] = [0.02, -0.01, 0.7]
    tf7, box7 = O.crop_window_tf(pose[None], K, 1.2, 0.2)
    assert box7[0][1] - box7[0][0] == 206
    _, got7 = O.warp_crop(rgb, xyz, tf7[0], box7[0])
    refn = cv2.warpPerspective(xyz, tf7[0].astype(np.float64), (160, 160), flags=cv2.INTER_NEAREST,
                               borderMode=cv2.BORDER_CONSTANT, borderValue=0)
    assert (np.abs(got7 - refn).max(-1) > 0).mean() < 0.03
    # window partly outside the image -> zeros (constant border)
    pose[:3, 3] = [-0.3, 0.0, 0.6]
    tf, box = O.crop_window_tf(pose[None], K, 1.2, 0.2)
    r2, x2 = O.warp_crop(rgb, xyz, tf[0], box[0])
    assert np.all(r2[:, :40] == 0) and np.all(x2[:, :40] == 0)


def test_normalize_pack_rules():
    pose = np.eye(4, dtype=F); pose[:3, 3] = [0.0, 0.0, 0.5]
    rgb = np.full((160, 160, 3), 0.25, F)
    xyz = np.zeros((160, 160, 3), F)
    xyz[0, 0] = [0.05, -0.05, 0.55]        # -> (0.5, -0.5, 0.5) in units of the radius 0.1
    xyz[0, 1] = [0.0, 0.0, 0.0005]         # z below min_depth -> all three zero
    xyz[0, 2] = [0.5, 0.0, 0.6]            # x = 5 radii > 4 -> x zeroed, others kept
    out = O.normalize_pack(rgb, xyz, pose, 0.2)
    np.testing.assert_allclose(out[0, 0], [0.25, 0.25, 0.25, 0.5, -0.5, 0.5], atol=1e-6)
    np.testing.assert_allclose(out[0, 1, 3:], 0)
    np.testing.assert_allclose(out[0, 2, 3:], [0, 0, 1.0], atol=1e-6)


def test_pose_update_known_answers():
    pose = np.eye(4, dtype=F)[None].copy()
    out = O.pose_update(pose, np.zeros((1, 3), F), np.zeros((1, 3), F), 0.2)
    np.testing.assert_array_equal(out, pose)          # zero delta -> identity (no NaN from normalising 0)
    out = O.pose_update(pose, np.array([[1, -2, 0.5]], F), np.zeros((1, 3), F), 0.2)
    np.testing.assert_allclose(out[0, :3, 3], [0.1, -0.2, 0.05], atol=1e-7)
    # saturated rotation about z: tanh(100) * 20 deg, applied transposed -> rotation by -20 deg
    out = O.pose_update(pose, np.zeros((1, 3), F), np.array([[0, 0, 100.0]], F), 0.2)
    a = math.radians(20)
    np.testing.assert_allclose(out[0, :3, :3], [[math.cos(a), math.sin(a), 0], [-math.sin(a), math.cos(a), 0],
                                                [0, 0, 1]], atol=1e-6)
    assert O.argmax_first([0.1, 0.7, 0.7, 0.2]) == 1  # first maximum wins


def test_oracle_render_agrees_with_independent_raycast():
    """The oracle renderer (raster + texture + shading + flip) at the ground-truth pose reproduces the
    analytically ray-cast scene crop (synth.make_scene shares no code with it)."""
    import fp_helpers as helpers
    from foundationpose_cpp_b200 import synth
    mesh = synth.uv_ellipsoid("m", slices=64, stacks=33, tex_size=256)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    xyz = O.depth_to_xyz(depth, synth.DEFAULT_K)
    ren, tra, info = O.render_and_transform(helpers.oracle_mesh(mesh), gt[None], rgb, xyz, synth.DEFAULT_K, 1.2)
    fg = (info["ids"][0][::-1] > 0)
    inner = fg & np.roll(fg, 3, 0) & np.roll(fg, -3, 0) & np.roll(fg, 3, 1) & np.roll(fg, -3, 1)
    assert inner.sum() > 2000
    assert np.median(np.abs(ren[0][inner][:, :3] - tra[0][inner][:, :3])) < 0.03     # colours
    valid = inner & (tra[0][..., 5] != 0)
    assert np.median(np.abs(ren[0][valid][:, 3:] - tra[0][valid][:, 3:])) < 0.03     # normalised xyz (3 mm / radius)


def test_clipper_agrees_with_double_precision_clipping():
    """Property: a triangle that crosses the near plane, rasterised through the restated cudaraster clipper
    (fp32, barycentric Sutherland-Hodgman), covers the same pixels as the polygon obtained by clipping it in
    float64 and fan-triangulating it by hand - up to a few boundary pixels along the cut, where the two
    paths round differently."""
    rng = np.random.default_rng(11)
    W = H = 64
    worst = 0
    drawn = 0
    for _ in range(40):
        # two vertices inside the depth range, one in front of the near plane (z/w < -1)
        v = np.zeros((3, 4), np.float64)
        v[:, 3] = rng.uniform(0.5, 2.0, 3)
        v[:, 0] = rng.uniform(-0.8, 0.8, 3) * v[:, 3]
        v[:, 1] = rng.uniform(-0.8, 0.8, 3) * v[:, 3]
        v[:2, 2] = rng.uniform(-0.8, 0.8, 2) * v[:2, 3]
        v[2, 2] = -rng.uniform(1.2, 3.0) * v[2, 3]
        ids, nclip = O.rasterize_ids(v.astype(F), np.array([[0, 1, 2]], np.int32), H, W)
        assert nclip == 1
        # float64 clip against w + z >= 0: the two crossing edges are 0-2 and 1-2
        d = v[:, 3] + v[:, 2]
        cut = [v[k] + (v[2] - v[k]) * (d[k] / (d[k] - d[2])) for k in (0, 1)]
        poly = np.array([v[0], v[1], cut[1], cut[0]])
        # pull the cut vertices a hair inside the range so that they take the fast path
        poly[2:, 2] = np.maximum(poly[2:, 2], -poly[2:, 3] * (1 - 1e-6))
        ids2, nclip2 = O.rasterize_ids(poly.astype(F), np.array([[0, 1, 2], [0, 2, 3]], np.int32), H, W)
        assert nclip2 == 0
        diff = int(((ids > 0) != (ids2 > 0)).sum())
        worst = max(worst, diff)
        drawn += int((ids > 0).sum())
        assert diff <= 0.02 * max(1, (ids2 > 0).sum()) + 3, diff
    assert drawn > 2000, "the random triangles must cover something"


def test_host_prepost_cpp_matches_the_numpy_oracle():
    """oracle/host_prepost.cpp (the reference's host-side functions in dependency-free C++, timed as part of the CPU
    baseline) against fp_oracle.py on a seeded scene: translation guess, crop transforms, bbox2d, clip matrices and
    the pose update must be the same numbers."""
    import ctypes
    from foundationpose_cpp_b200 import synth
    lib = O.host_prepost_lib()
    mesh = synth.uv_ellipsoid("m", slices=16, stacks=9, tex_size=32, seed=2)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    K = synth.DEFAULT_K
    p = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    Kf = np.ascontiguousarray(np.asarray(K, np.float32).reshape(9))
    c = np.zeros(3, np.float32)
    assert lib.host_guess_translation(p(np.ascontiguousarray(depth, np.float32)), p(np.ascontiguousarray(mask)), 480, 640, p(Kf), p(c)) == 0
    np.testing.assert_allclose(c, O.guess_translation(depth, mask, K), rtol=1e-6)
    assert lib.host_guess_translation(p(np.ascontiguousarray(depth, np.float32)), p(np.zeros_like(mask)), 480, 640, p(Kf), p(c)) != 0
    poses = O.rotation_grid()[::17].copy()
    poses[:, :3, 3] = c
    n = len(poses)
    flat = np.ascontiguousarray(poses.reshape(n, 16), np.float32)
    tf, bbox, mvp = np.zeros((n, 9), np.float32), np.zeros((n, 4), np.float32), np.zeros((n, 16), np.float32)
    lib.host_pose_setup(p(flat), n, p(Kf), 480, 640, ctypes.c_float(mesh.diameter), ctypes.c_float(1.2), p(tf), p(bbox), p(mvp))
    ref_tf, _ = O.crop_window_tf(poses, K, 1.2, mesh.diameter)
    np.testing.assert_array_equal(tf.reshape(n, 3, 3), ref_tf)
    np.testing.assert_array_equal(bbox, O.bbox2d_from_tf(ref_tf))
    proj = O.projection_matrix(K, 480, 640)
    for i in range(n):
        np.testing.assert_array_equal(mvp[i].reshape(4, 4), O._mat4_mul(proj, O._mat4_mul(O.GL_CAM_IN_CV_CAM, poses[i])))
    rng = np.random.default_rng(0)
    trans, rot = rng.normal(0, 0.2, (n, 3)).astype(np.float32), rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    upd = flat.copy()
    lib.host_refine_post(p(upd), n, p(trans), p(rot), ctypes.c_float(mesh.diameter))
    np.testing.assert_allclose(upd.reshape(n, 4, 4), O.pose_update(poses, trans, rot, mesh.diameter), atol=2e-6)
    ms1 = O.host_register_work_ms(depth, mask, K, O.rotation_grid(), mesh.diameter, threads=1, repeat=3)
    assert 0 < ms1 < 1000
