import os, sys, tempfile, json, time
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from foundationpose_cpp_b200 import api, synth, weights
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
kind = sys.argv[3] if len(sys.argv) > 3 else "mesh10k"
d = tempfile.mkdtemp()
rp, sp = weights.write_seeded(d)
if kind == "mesh10k":
    mesh = synth.uv_ellipsoid("mesh10k", slices=100, stacks=51, axes=(0.08, 0.08, 0.1), tex_size=1024, seed=0)
else:
    mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=0)
rgb, depth, mask, gt = synth.make_scene(mesh)
m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480, max_w=640, max_hyps=n)
for i in range(3):
    ok, pose = m.Register(rgb, depth, mask, mesh.name, refine_itr=1)
    assert ok
m.set_async(True)
for i in range(3):
    m.register_async(mesh.name, 1)
m.wait()
st = torch.cuda.ExternalStream(m.stream_ptr())
best = 1e9
for trial in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for i in range(reps):
        m.register_async(mesh.name, 1)
    e1.record(st)
    m.wait()
    best = min(best, e0.elapsed_time(e1) / reps)
print(json.dumps({"n": n, "device_ms_per_register": round(best, 4), "launches": m.last_launch_count()}))
