"""One small Register (N hypotheses) + one Track through the public API, for compute-sanitizer runs:
    compute-sanitizer --tool racecheck python tools/sanitize_register.py 4
    compute-sanitizer --tool memcheck  python tools/sanitize_register.py 4
(CUDA graphs are off so that every kernel is attributed to its own launch.)"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from foundationpose_cpp_b200 import api, synth, weights  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    rp, sp = weights.write_seeded(tempfile.mkdtemp())
    mesh = synth.uv_ellipsoid("obj", slices=48, stacks=25, tex_size=128, seed=0)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    m = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], synth.DEFAULT_K, max_h=480,
                                      max_w=640, max_hyps=n, use_cuda_graph=False)
    ok, pose = m.Register(rgb, depth, mask, "obj", refine_itr=1)
    assert ok, m.last_error
    ok, pose2 = m.Track(rgb, depth, pose, "obj", refine_itr=1)
    assert ok, m.last_error
    print(f"sanitize_register: Register (N = {n}) and Track completed, z = {pose[2, 3]:.4f} / {pose2[2, 3]:.4f}")
    m.close()


if __name__ == "__main__":
    main()
