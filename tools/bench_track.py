"""BASELINE config 3: Track (N=1, refine-only) fps with p50 / p99 latency through the public call
(host buffers, H2D + D2H + sync inside), like the reference's speed_track loop
(simple_tests/src/test_foundationpose.cpp:130-155)."""
import json
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from foundationpose_cpp_b200 import api, synth, weights  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    refine_itr = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=0)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    rp, sp = weights.write_seeded(tempfile.mkdtemp(prefix="fpb200_w_"))
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, 480, 640)
    rgb_p = torch.from_numpy(rgb).pin_memory().numpy()
    depth_p = torch.from_numpy(depth).pin_memory().numpy()
    pose = synth.perturb_pose(gt, 5.0, 0.01, 3)
    for _ in range(20):
        ok, out = model.Track(rgb_p, depth_p, pose, "mustard_proxy", refine_itr)
        assert ok
    lat = np.zeros(calls)
    t0 = time.perf_counter()
    for i in range(calls):
        t1 = time.perf_counter()
        ok, out = model.Track(rgb_p, depth_p, pose, "mustard_proxy", refine_itr)
        lat[i] = time.perf_counter() - t1
    dt = time.perf_counter() - t0
    # several objects tracked from one host thread: S sessions, enqueue all, then collect all
    sessions = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    models = [model] + [api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)],
                                                      synth.DEFAULT_K, 480, 640, max_hyps=1) for _ in range(sessions - 1)]
    for _ in range(20):
        for m in models:
            m.track_async(rgb_p, depth_p, pose, "mustard_proxy", refine_itr)
        outs = [m.track_wait() for m in models]
    assert all(np.array_equal(o, out) for o in outs)
    rounds = max(1, calls // sessions)
    t2 = time.perf_counter()
    for _ in range(rounds):
        for m in models:
            m.track_async(rgb_p, depth_p, pose, "mustard_proxy", refine_itr)
        for m in models:
            m.track_wait()
    dt_multi = time.perf_counter() - t2
    print(json.dumps({"metric": "Track fps (N=1, refine-only)", "value": calls / dt,
                      "multi_session": {"sessions": sessions, "fps_aggregate": rounds * sessions / dt_multi}, "unit": "fps", "calls": calls,
                      "refine_itr": refine_itr, "p50_ms": float(np.percentile(lat, 50) * 1e3),
                      "p99_ms": float(np.percentile(lat, 99) * 1e3), "launches_per_call": model.last_launch_count(),
                      "published_reference_rtx4060_fps": 220}))
    model.close()


if __name__ == "__main__":
    main()
