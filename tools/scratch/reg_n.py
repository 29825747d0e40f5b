import os, sys, tempfile, json, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from foundationpose_cpp_b200 import api, synth, weights
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
kind = sys.argv[3] if len(sys.argv) > 3 else "mesh10k"
d = tempfile.mkdtemp()
rp, sp = weights.write_seeded(d)
if kind == "mesh10k":
    mesh = synth.uv_ellipsoid("mesh10k", slices=100, stacks=51, axes=(0.08, 0.08, 0.1), tex_size=1024, seed=0)
else:
    mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=0)
rgb, depth, mask, gt = synth.make_scene(mesh)
m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480, max_w=640, max_hyps=n)
for i in range(reps):
    t0 = time.perf_counter()
    ok, pose = m.Register(rgb, depth, mask, mesh.name, refine_itr=1)
    dt = time.perf_counter() - t0
    assert ok
print(json.dumps({"n": n, "last_call_ms": dt * 1e3, "launches": m.last_launch_count()}))
