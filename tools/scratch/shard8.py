import os, sys, time, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import fp_helpers as helpers
from foundationpose_cpp_b200 import api, synth
N = 252
world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
d = tempfile.mkdtemp()
rp, sp = helpers.write_seeded_weights(d)
mesh = synth.uv_ellipsoid("obj", slices=64, stacks=33, tex_size=256, seed=0)
rgb, depth, mask, _ = synth.make_scene(mesh)
def make(rank, world):
    return api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K,
                                         max_h=480, max_w=640, max_hyps=N, shard_rank=rank, shard_world=world, use_cuda_graph=(os.environ.get("NOGRAPH") is None))
ms = [make(r, world) for r in range(world)]
api.connect_local(ms)
for m in ms:
    m.upload(rgb, depth); m.sample_hypotheses(mask, "obj")
if os.environ.get("WARM"):
    for m in ms:
        m.refine_step(); m.score_features()
    torch.cuda.synchronize()
for rep in range(4):
    ts = []
    for m in ms:
        t0 = time.perf_counter(); m.register_async("obj", 1); ts.append(time.perf_counter() - t0)
    res = []
    for m in ms:
        try:
            idx, pose = m.wait(); res.append(idx)
        except Exception as e:
            res.append(str(e)[:40])
    if rep == 0:
        for i, m in enumerate(ms):
            print("  session", i, "stage ms", [round(x, 2) for x in m.last_timings()], flush=True)
    print("rep", rep, "enqueue ms", [round(t * 1e3, 1) for t in ts], "res", res, flush=True)
