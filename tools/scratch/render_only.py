import os, sys, tempfile, json
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from foundationpose_cpp_b200 import api, synth, weights
d = tempfile.mkdtemp()
rp, sp = weights.write_seeded(d)
mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=0)
rgb, depth, mask, gt = synth.make_scene(mesh)
m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480, max_w=640)
m.upload(rgb, depth); m.sample_hypotheses(mask, mesh.name)
print(m.time_stage(0, 3))
