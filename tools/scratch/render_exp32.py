import os, sys, tempfile, json
ROOT = "/root/repo"
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from foundationpose_cpp_b200 import api, synth, weights
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
d = tempfile.mkdtemp()
rp, sp = weights.write_seeded(d)
mesh = synth.uv_ellipsoid("mesh10k", slices=100, stacks=51, axes=(0.08, 0.08, 0.1), tex_size=1024, seed=0)
rgb, depth, mask, gt = synth.make_scene(mesh)
m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480, max_w=640, max_hyps=n)
m.upload(rgb, depth); m.sample_hypotheses(mask, mesh.name)
print(json.dumps({"n": n, "slices": os.environ.get("FPB200_SPLIT_SLICES"), "rows": os.environ.get("FPB200_SPLIT_SHADE_ROWS"), "render_ms": round(m.time_stage(0, 200), 4)}))
