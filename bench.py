"""Headline benchmark: Register pose-hypotheses/sec (N=252 hypotheses, 160x160 crops, refine_itr=1).

  python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W     # CPU restatement of the reference path

A step = one Register over 252 hypotheses of the seeded "mustard-proxy" scene (no dataset is
reachable offline): sampler -> (render + crop/warp -> refine net -> pose update) x refine_itr ->
render + crop/warp -> score net -> cross-hypothesis tail -> argmax.
  value : device-timed (CUDA events on the session stream), rgb/depth/mask already resident in HBM
  e2e   : the public Register call with pinned HOST buffers: H2D copies, D2H of the pose, syncs included
With --gpus N > 1 (torchrun) the 252 hypotheses are sharded over the ranks; the only exchange is one
all-gather of [528]-float rows (512-d scorer feature + refined pose) per hypothesis before the tail.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HYP = 252
METRIC = "Register pose-hypotheses/sec (N=252, 160x160)"
UNIT = "hyp/s"
WORKLOAD = "mustard-proxy Register: 8100-tri textured mesh, 640x480 RGB-D, N=252 hypotheses, refine_itr=1"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refine-itr", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="shard", choices=["shard", "replicas"],
                    help="N > 1: shard the 252 hypotheses of one Register over the ranks (BASELINE config 4, strong "
                         "scaling) or run one independent object per GPU (config 5, weak scaling, no collective)")
    return ap.parse_args()


def scene():
    from foundationpose_cpp_b200 import synth
    mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=0)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    return mesh, rgb, depth, mask, gt, synth.DEFAULT_K


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        self.marks = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [(ts, line) for ts, line in self.rows if t0 - 0.05 <= ts <= t1 + 0.05]
        if not inside:  # timed region shorter than the sampling period: take the samples closest to it
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
        for ts, line in inside:
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
                pw.append(float(f[2]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_reference_step(ctx, n_hyp: int):
    """One bounded sample of the workload on the host: n_hyp hypotheses through the whole path
    (oracle/ restatement of the reference: C rasteriser + numpy + PyTorch-CPU networks)."""
    import numpy as np
    import torch
    O, nets = ctx["O"], ctx["nets"]
    poses = ctx["poses"][:n_hyp]
    xyz = O.depth_to_xyz(ctx["depth"], ctx["K"])
    ren, tra, _ = O.render_and_transform(ctx["omesh"], poses, ctx["rgb"], xyz, ctx["K"], O.REFINE_CROP_RATIO)
    with torch.no_grad():
        t, r = ctx["refiner"](nets.nhwc_to_inputs(torch.from_numpy(ren)), nets.nhwc_to_inputs(torch.from_numpy(tra)))
    poses = O.pose_update(poses, t.numpy(), r.numpy(), ctx["omesh"]["diameter"])
    ren, tra, _ = O.render_and_transform(ctx["omesh"], poses, ctx["rgb"], xyz, ctx["K"], O.SCORE_CROP_RATIO)
    with torch.no_grad():
        s = ctx["scorer"](nets.nhwc_to_inputs(torch.from_numpy(ren)), nets.nhwc_to_inputs(torch.from_numpy(tra)))
    return poses[O.argmax_first(s.numpy())]


def cpu_context():
    import numpy as np
    import torch
    from foundationpose_cpp_b200 import weights
    from oracle import fp_oracle as O
    from oracle import nets
    mesh, rgb, depth, mask, gt, K = scene()
    tc = mesh.texcoords.astype(np.float32).copy()
    tc[:, 1] = np.float32(1) - tc[:, 1]
    omesh = {"vertices": (mesh.vertices - mesh.center).astype(np.float32), "normals": mesh.normals, "texcoords": tc,
             "faces": mesh.faces, "texture": (mesh.texture.astype(np.float32) * np.float32(1.0 / 255.0)),
             "diameter": float(mesh.diameter)}
    refiner, scorer = nets.RefineNet(), nets.ScoreNet()
    refiner.load_state_dict(weights.seeded_state_dict(weights.KIND_REFINER, 2))
    scorer.load_state_dict(weights.seeded_state_dict(weights.KIND_SCORER, 3))
    refiner.eval(), scorer.eval()
    poses = O.sample_hypotheses(depth, mask, K)
    return {"O": O, "nets": nets, "omesh": omesh, "rgb": rgb, "depth": depth, "K": K, "poses": poses,
            "refiner": refiner, "scorer": scorer, "cores": torch.get_num_threads()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ctx = cpu_context()
    per_step = 8  # hypotheses per step: ~370 GFLOP of fp32 network + raster/warp, about a second on the host
    for _ in range(max(args.warmup, 1)):
        cpu_reference_step(ctx, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_reference_step(ctx, per_step)
    dt = time.perf_counter() - t0
    value = args.steps * per_step / dt
    sample = f"{per_step} of 252 hypotheses per step through the full path (sampler excluded), fp32, all host threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "CPU restatement (oracle/) of the reference path; the reference's "
                   "TensorRT/CV-CUDA build cannot be compiled in this image"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ctx["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- CUDA arm
class _DevMem:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from foundationpose_cpp_b200 import api, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    replicas = args.mode == "replicas" and world > 1
    sw, sr = (1, 0) if replicas else (world, rank)  # shard world / rank seen by the session
    mesh, rgb, depth, mask, gt, K = scene()
    wdir = tempfile.mkdtemp(prefix=f"fpb200_w{rank}_")
    rp, sp = weights.write_seeded(wdir)
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_h=480, max_w=640,
                                          device=local_rank, max_hyps=N_HYP, shard_rank=sr, shard_world=sw)
    # pinned host copies of the frame (what a capture pipeline would hand over)
    rgb_p = torch.from_numpy(rgb).pin_memory().numpy()
    depth_p = torch.from_numpy(depth).pin_memory().numpy()
    mask_p = torch.from_numpy(mask).pin_memory().numpy()
    stream = torch.cuda.ExternalStream(model.stream_ptr(), device=torch.device("cuda", local_rank))
    name = "mustard_proxy"
    from foundationpose_cpp_b200 import sharding
    cap = sharding.shard_capacity(N_HYP, sw)

    def gather():
        if sw == 1:
            return
        lp, nl, off, ap, na = model.feature_buffers()
        src = torch.as_tensor(_DevMem(lp, cap * 528), device=f"cuda:{local_rank}")
        dst = torch.as_tensor(_DevMem(ap, world * cap * 528), device=f"cuda:{local_rank}")
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(dst, src)

    def step_resident():
        if sw == 1 and not profiling[0]:
            model.register_async(name, args.refine_itr)  # one CUDA-graph replay per Register
            return
        model.sample_hypotheses(None, name)
        for _ in range(args.refine_itr):
            model.refine_step()
        model.score_features()
        gather()
        model.score_finish_async()

    profiling = [False]

    def step_e2e():
        if sw == 1:
            ok, pose = model.Register(rgb_p, depth_p, mask_p, name, refine_itr=args.refine_itr)
            assert ok, model.last_error
            return pose
        model.upload(rgb_p, depth_p)
        model.sample_hypotheses(mask_p, name)
        for _ in range(args.refine_itr):
            model.refine_step()
        model.score_features()
        gather()
        model.score_finish_async()
        return model.wait()[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=f"cuda:{local_rank}", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also makes rgb/depth/mask resident); the clock sampler starts now so that it is already
    # streaming when the (possibly very short) timed region begins
    clocks = ClockSampler(local_rank)
    model.set_async(sw > 1)
    for _ in range(max(args.warmup, 3)):
        pose_w = step_e2e()
    model.set_async(True)
    for _ in range(2):
        step_resident()
    model.wait()

    # ---- timed region 1: device-resident inputs, CUDA events on the launching stream
    launches0 = model.launch_count_total()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_w0 = time.time()
    ev0.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev1.record(stream)
    idx, pose = model.wait()
    barrier()
    dev_ms = max_over_ranks(ev0.elapsed_time(ev1))
    launches = model.launch_count_total() - launches0
    units = N_HYP * (world if replicas else 1)  # hypotheses processed per step by all ranks
    value = args.steps * units / (dev_ms * 1e-3)

    # ---- timed region 1b: the same K steps with every conv_gemm launch bracketed by CUDA events on
    # the session stream (the brackets cost a few % of step time, so `value` comes from the clean pass)
    model.set_gemm_profile(True)
    profiling[0] = True
    model.gemm_profile()  # clear
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(args.steps):
        step_resident()
    ev3.record(stream)
    model.wait()
    barrier()
    prof_ms = ev2.elapsed_time(ev3)
    gemm_ms, gemm_flops, gemm_launches = model.gemm_profile()
    model.set_gemm_profile(False)
    profiling[0] = False

    # ---- timed region 2: public call with host buffers (H2D + D2H + syncs inside)
    model.set_async(sw > 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pose_e = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    t_w2 = time.time()
    clk = clocks.stop(t_w0, t_w2)
    e2e_value = args.steps * units / e2e_s

    f_ref, f_sco = model.flops_per_hypothesis()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (measured)" if peaks else "fallback 1.4 PFLOP/s sustained"
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "conv_gemm_traffic.json"))).get("dram_bytes_per_launch")
    except Exception:
        pass
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak" if replicas else "strong",
        "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOAD, "hypotheses": N_HYP, "refine_itr": args.refine_itr,
                   "sharding": ("single GPU" if world == 1 else
                                f"{world} independent objects, one per GPU, no collective" if replicas else
                                f"hypotheses split over {world} rank(s); one all-gather of [N,528] f32 per Register"),
                   "l2": "no explicit flush: each step streams ~2.6 GB of activations through HBM (>> 126 MB L2); "
                         "weights (66 MB) and the frame (2.4 MB) are legitimately hot",
                   "weights": "seeded random init of the published architecture (no checkpoint offline)",
                   "flops_per_hypothesis": {"refiner": f_ref, "scorer": f_sco},
                   "whole_job_tflops": value * (f_ref * args.refine_itr + f_sco) / 1e12,
                   "published_reference_rtx4060_hyp_s": 705.6},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rgb.nbytes + depth.nbytes + mask.nbytes),
                "d2h_bytes_per_step": 72, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM; all conv / linear layers)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src, "launches_timed": int(gemm_launches),
                     "kernel_ms_per_step": gemm_ms / args.steps, "kernel_share_of_step": gemm_ms / prof_ms if prof_ms else None,
                     "timing": "second pass of the same K steps with CUDA events around every conv_gemm launch"},
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        ctx = cpu_context()
        cpu_reference_step(ctx, 1)
        n_s, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            cpu_reference_step(ctx, 2)
            n_s += 2
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n_s / dt, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                               "sample": f"{n_s} hypotheses (2 per call) through the full oracle path in {dt:.1f} s"}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
