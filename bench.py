"""Headline benchmark: Register pose-hypotheses/sec (N=252 hypotheses, 160x160 crops, refine_itr=1).

  python bench.py --gpus N --steps K --warmup W            # this repository (CUDA, sm_100a)
  python bench.py --impl reference --steps K --warmup W     # CPU restatement of the reference path

A step = one Register over 252 hypotheses (no dataset is reachable offline, scenes are seeded synthetic):
sampler -> (render + crop/warp -> refine net -> pose update) x refine_itr -> render + crop/warp -> score net ->
cross-hypothesis tail -> argmax.
  value : device-timed (CUDA events on the session stream), rgb / depth / mask already resident in HBM; every step is
          one CUDA-graph replay enqueued by fpb200_register_async
  e2e   : the public Register call with pinned HOST buffers: H2D copies, D2H of the pose, syncs included
N = 1 runs BASELINE config 2 (mustard-proxy scene).  N > 1 (torchrun) runs BASELINE config 4: the 10 000-triangle mesh,
the 252 hypotheses sharded over the ranks; the only exchange (512-d scorer feature + refined pose per hypothesis)
happens INSIDE the library over NVLink peer memory (fpb200_exchange_*), so a sharded Register is still one graph
replay per rank.  The other BASELINE configs are reported as sub-records under "configs" of the same JSON line
(config 1 N=32, config 3 Track, config 4 on one GPU, config 5 one object per GPU).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes as C
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
WORKLOADS = {
    "mustard": "mustard-proxy Register: 8100-tri textured mesh, 640x480 RGB-D, N=252 hypotheses, refine_itr=1",
    "mesh10k": "synthetic 10k-tri mesh Register: 10000-tri textured mesh, 640x480 RGB-D, N=252 hypotheses, refine_itr=1",
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--refine-itr", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records of the other BASELINE configs")
    ap.add_argument("--workload", default=None, choices=[None, "mustard", "mesh10k"],
                    help="default: mustard on one GPU (config 2), mesh10k on several (config 4)")
    ap.add_argument("--mode", default="shard", choices=["shard", "replicas"],
                    help="N > 1: shard the 252 hypotheses of one Register over the ranks (BASELINE config 4, strong "
                         "scaling) or run one independent object per GPU (config 5, weak scaling, no exchange)")
    ap.add_argument("--host-gather", action="store_true",
                    help="N > 1: exchange through a host-driven NCCL all-gather between stage calls (the round-1 path) "
                         "instead of the in-library peer-memory exchange")
    return ap.parse_args()


def scene(kind: str = "mustard", seed: int = 0):
    from foundationpose_cpp_b200 import synth
    if kind == "mesh10k":  # BASELINE config 4: UV sphere 100 x 51 -> exactly 10 000 triangles
        mesh = synth.uv_ellipsoid(f"mesh10k_{seed}", slices=100, stacks=51, axes=(0.08, 0.08, 0.1), tex_size=1024, seed=seed)
    else:
        mesh = synth.uv_ellipsoid("mustard_proxy", slices=90, stacks=46, axes=(0.045, 0.045, 0.095), tex_size=512, seed=seed)
    rgb, depth, mask, gt = synth.make_scene(mesh)
    return mesh, rgb, depth, mask, gt, synth.DEFAULT_K


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons of one GPU, sampled DURING the timed region: NVML in-process (a thread polling
    every 5 ms: works for timed regions of a few tens of ms and with 8 ranks starting at once), `nvidia-smi -lms 20`
    as a fallback when NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int, uuid: str = None):
        self.rows, self.proc, self.nvml, self.stop_flag, self.mx = [], None, None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            if uuid:
                for cand in (uuid, "GPU-" + uuid):
                    try:
                        h = pynvml.nvmlDeviceGetHandleByUUID(cand if isinstance(cand, bytes) else cand.encode())
                        break
                    except Exception:
                        h = None
            if h is None:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
                h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml, self.handle = pynvml, h
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _poll(self):
        n = self.nvml
        bits = [(n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown") else n.nvmlClocksThrottleReasonHwSlowdown, "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)), "sw_power_cap")]
        while not self.stop_flag:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                mask = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                self.rows.append((time.time(), sm, pw, [nm for b, nm in bits if mask & b]))
            except Exception:
                pass
            time.sleep(0.005)

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                self.mx = float(f[1])
                self.rows.append((time.time(), float(f[0]), float(f[2]),
                                  [nm for nm, v in zip(self.NAMES, f[3:7]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self, t0: float, t1: float):
        if not self.proc and not self.nvml:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi / NVML unavailable"]}
        time.sleep(0.05)
        self.stop_flag = True
        if self.proc:
            time.sleep(0.1)
            self.proc.terminate()
        inside = [r for r in self.rows if t0 <= r[0] <= t1]
        if not inside:  # timed region shorter than the sampling period: take the samples closest to it
            inside = sorted(self.rows, key=lambda r: min(abs(r[0] - t0), abs(r[0] - t1)))[:3]
        sm = sorted(r[1] for r in inside)
        pw = sorted(r[2] for r in inside)
        reasons = sorted({nm for r in inside for nm in r[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None,
                "source": "NVML in-process, 5 ms period" if self.nvml else "nvidia-smi -lms 20"}


# ----------------------------------------------------------------------------- CPU arm (oracle port)
def cpu_context(kind: str = "mustard"):
    import numpy as np
    import torch
    from foundationpose_cpp_b200 import weights
    from oracle import fp_oracle as O
    from oracle import nets
    mesh, rgb, depth, mask, gt, K = scene(kind)
    tc = mesh.texcoords.astype(np.float32).copy()
    tc[:, 1] = np.float32(1) - tc[:, 1]
    omesh = {"vertices": (mesh.vertices - mesh.center).astype(np.float32), "normals": mesh.normals, "texcoords": tc,
             "faces": mesh.faces, "texture": (mesh.texture.astype(np.float32) * np.float32(1.0 / 255.0)),
             "diameter": float(mesh.diameter)}
    refiner, scorer = nets.RefineNet(), nets.ScoreNet()
    refiner.load_state_dict(weights.seeded_state_dict(weights.KIND_REFINER, 2))
    scorer.load_state_dict(weights.seeded_state_dict(weights.KIND_SCORER, 3))
    refiner.eval(), scorer.eval()
    return {"O": O, "nets": nets, "omesh": omesh, "rgb": rgb, "depth": depth, "mask": mask, "K": K,
            "refiner": refiner, "scorer": scorer, "cores": torch.get_num_threads()}


def cpu_reference_step(ctx, n_hyp: int, refine_itr: int = 1):
    """One bounded sample of the workload on the host (oracle/ restatement of the reference: numpy + C rasteriser +
    PyTorch-CPU networks): the WHOLE sampler, then n_hyp of the 252 hypotheses through refine / score / argmax.
    Returns (seconds in the sampler, seconds in the per-hypothesis part)."""
    import torch
    O, nets = ctx["O"], ctx["nets"]
    t0 = time.perf_counter()
    poses = O.sample_hypotheses(ctx["depth"], ctx["mask"], ctx["K"])[:n_hyp]
    xyz = O.depth_to_xyz(ctx["depth"], ctx["K"])
    t1 = time.perf_counter()
    for _ in range(refine_itr):
        ren, tra, _ = O.render_and_transform(ctx["omesh"], poses, ctx["rgb"], xyz, ctx["K"], O.REFINE_CROP_RATIO)
        with torch.no_grad():
            t, r = ctx["refiner"](nets.nhwc_to_inputs(torch.from_numpy(ren)), nets.nhwc_to_inputs(torch.from_numpy(tra)))
        poses = O.pose_update(poses, t.numpy(), r.numpy(), ctx["omesh"]["diameter"])
    ren, tra, _ = O.render_and_transform(ctx["omesh"], poses, ctx["rgb"], xyz, ctx["K"], O.SCORE_CROP_RATIO)
    with torch.no_grad():
        s = ctx["scorer"](nets.nhwc_to_inputs(torch.from_numpy(ren)), nets.nhwc_to_inputs(torch.from_numpy(tra)))
    _ = poses[O.argmax_first(s.numpy())]
    return t1 - t0, time.perf_counter() - t1


def host_prepost_baseline(ctx):
    """the reference's host-side functions (GuessTranslation, crop windows, bbox2d, clip
    matrices, RefinePostProcess) in dependency-free C++ (oracle/host_prepost.cpp), one Register's worth, timed on one
    thread and on all host threads."""
    O = ctx["O"]
    try:
        cores = os.cpu_count() or 1
        filt = O.bilateral_filter_depth(O.erode_depth(ctx["depth"]))
        grid = O.rotation_grid()
        ms1 = O.host_register_work_ms(filt, ctx["mask"], ctx["K"], grid, ctx["omesh"]["diameter"], threads=1, repeat=20)
        msn = O.host_register_work_ms(filt, ctx["mask"], ctx["K"], grid, ctx["omesh"]["diameter"], threads=cores, repeat=20)
        return {"ms_1_thread": ms1, "ms_all_threads": msn, "threads": cores,
                "what": "GuessTranslation + 2 x 252 x (ComputeCropWindowTF, ConstructBBox2D, Proj*GL*pose) + 252 x "
                        "RefinePostProcess in plain C++ (oracle/host_prepost.cpp); the product does all of it on the device"}
    except Exception as e:  # pragma: no cover
        return {"unavailable": str(e)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    kind = args.workload or ("mustard" if args.gpus <= 1 else "mesh10k")
    ctx = cpu_context(kind)
    for _ in range(max(min(args.warmup, 2), 1)):
        cpu_reference_step(ctx, 1, args.refine_itr)
    # size the per-step sample so that the whole run stays within ~2 minutes
    ts, th = cpu_reference_step(ctx, 2, args.refine_itr)
    per_hyp = th / 2
    budget = min(110.0 / max(args.steps, 1), 25.0) - ts
    per_step = int(max(1, min(N_HYP, budget / max(per_hyp, 1e-3))))
    t0 = time.perf_counter()
    t_s = t_h = 0.0
    for _ in range(args.steps):
        a, b = cpu_reference_step(ctx, per_step, args.refine_itr)
        t_s += a
        t_h += b
    dt = time.perf_counter() - t0
    # rate of the FULL 252-hypothesis step: the sampler runs once per Register, the rest scales with the hypotheses
    full_step_s = t_s / args.steps + N_HYP * (t_h / (args.steps * per_step))
    value = N_HYP / full_step_s
    sample = (f"each step = the whole sampler + {per_step} of the 252 hypotheses through refine / score / argmax (fp32, "
              f"{ctx['cores']} host threads); value = 252 / (sampler + 252 x per-hypothesis time), i.e. the full step "
              f"extrapolated from the sample")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[kind], "hypotheses": N_HYP, "refine_itr": args.refine_itr,
                   "sample": sample,
                   "note": "CPU restatement (oracle/) of the reference path; the reference's TensorRT / CV-CUDA build "
                           "cannot be compiled in this image (no TensorRT, CV-CUDA, Eigen, OpenCV-C++, assimp, glog)",
                   "full_step_ms_extrapolated": full_step_s * 1e3},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": ctx["cores"], "kind": "port", "sample": sample,
                         "host_pre_post": host_prepost_baseline(ctx)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------- CUDA arm helpers
class _DevMem:
    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}


def _pin(a):
    import torch
    return torch.from_numpy(a).pin_memory().numpy()


def timed_async_registers(model, name, steps, refine_itr, stream, torch):
    """K graph-replayed Registers on the resident frame between two CUDA events on the session stream -> ms."""
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        model.register_async(name, refine_itr)
    ev1.record(stream)
    model.wait()
    return ev0.elapsed_time(ev1)


def extra_single_gpu(args, model, mesh, rgb_p, depth_p, mask_p, gt, K, rp, sp, local_rank, torch, np):
    """Sub-records of the other BASELINE configs on one GPU."""
    from foundationpose_cpp_b200 import api, synth
    out = {}
    name = mesh.name
    dev = torch.device("cuda", local_rank)
    # ---- per-stage device times of one eager Register (CUDA events between the stages)
    try:
        m2 = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_h=480, max_w=640,
                                           device=local_rank, max_hyps=N_HYP, use_cuda_graph=False)
        for _ in range(3):
            ok, _ = m2.Register(rgb_p, depth_p, mask_p, name, refine_itr=args.refine_itr)
        t = m2.last_timings()
        out["stage_ms_eager"] = {"upload+xyz": t[0], "sampler": t[1], "refine (render, warp, net, update)": t[3],
                                 "score (render, warp, net)": t[4], "tail + argmax": t[5]}
        # the render / warp kernels alone, against the HBM roofline (algorithmic bytes: the packed fp16 crops written)
        m2.upload(rgb_p, depth_p)
        m2.sample_hypotheses(mask_p, name)
        r_ms, w_ms = m2.time_stage(0, 20), m2.time_stage(1, 20)
        bytes_crop = N_HYP * 160 * 160 * 8 * 2
        out["preprocess_kernels"] = {"render_ms": r_ms, "warp_ms": w_ms, "algorithmic_bytes_each": bytes_crop,
                                     "render_gbs": bytes_crop / (r_ms * 1e-3) / 1e9, "warp_gbs": bytes_crop / (w_ms * 1e-3) / 1e9,
                                     "what": "RENDER = pose set-up + vertex stage + rasterise + shade + normalise + pack of 252 "
                                             "hypotheses; WARP = observed crop + normalise + pack"}
        m2.close()
    except Exception as e:  # pragma: no cover
        out["stage_ms_eager"] = {"unavailable": str(e)}
    # ---- same-box reference CUDA preprocess: the reference's own render kernels, N = 252
    try:
        out["reference_cuda_preprocess"] = reference_cuda_preprocess(mesh, rgb_p, depth_p, mask_p, K, np)
    except Exception as e:
        out["reference_cuda_preprocess"] = {"unavailable": str(e)}
    # ---- config 1: N = 32 hypotheses
    try:
        m1 = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_h=480, max_w=640,
                                           device=local_rank, max_hyps=32)
        s1 = torch.cuda.ExternalStream(m1.stream_ptr(), device=dev)
        for _ in range(3):
            ok, _ = m1.Register(rgb_p, depth_p, mask_p, name, refine_itr=args.refine_itr)
        m1.set_async(True)
        ms = timed_async_registers(m1, name, args.steps, args.refine_itr, s1, torch)
        m1.set_async(False)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            m1.Register(rgb_p, depth_p, mask_p, name, refine_itr=args.refine_itr)
        e2e = time.perf_counter() - t0
        out["config1_register_n32"] = {"value": 32 * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_register": ms / args.steps,
                                       "e2e_ms_per_register": e2e / args.steps * 1e3, "dtype": "f16",
                                       "note": "BASELINE config 1 asks for fp32; this library computes in fp16 with fp32 "
                                               "accumulation (what the reference's --fp16 engines do); parity of this "
                                               "configuration against the fp32 oracle pipeline: "
                                               "tests/test_bench_scale_parity_gpu.py"}
        m1.close()
    except Exception as e:  # pragma: no cover
        out["config1_register_n32"] = {"unavailable": str(e)}
    # ---- config 3: Track (N = 1, refine only), the reference's speed_track protocol: consecutive public calls
    try:
        calls = 5000
        pose = synth.perturb_pose(gt, 5.0, 0.01, 3)
        for _ in range(20):
            ok, o = model.Track(rgb_p, depth_p, pose, name, 1)
            assert ok
        lat = np.zeros(calls)
        t0 = time.perf_counter()
        for i in range(calls):
            t1 = time.perf_counter()
            model.Track(rgb_p, depth_p, pose, name, 1)
            lat[i] = time.perf_counter() - t1
        dt = time.perf_counter() - t0
        out["config3_track"] = {"value": calls / dt, "unit": "fps", "calls": calls, "p50_ms": float(np.percentile(lat, 50) * 1e3),
                                "p99_ms": float(np.percentile(lat, 99) * 1e3), "launches_per_call": model.last_launch_count(),
                                "published_reference_rtx4060_fps": 220,
                                "protocol": "simple_tests/src/test_foundationpose.cpp:130-155 (consecutive Track calls, host buffers)"}
        # the same frames as a stream (fpb200_track_stream_*): upload of frame k + 1 on a copy stream while frame k is
        # refined, pose chained in device memory, two frames in flight (host buffers, results read back every frame)
        try:
            rgb_q, depth_q = _pin(rgb_p.copy()), _pin(depth_p.copy())  # two pinned host frames, alternating
            hostf = [(rgb_p, depth_p), (rgb_q, depth_q)]
            for rep in range(2):
                n_s = 40 if rep == 0 else calls
                model.track_stream_begin(pose, name, 1)
                model.track_stream_push(*hostf[0])
                t0 = time.perf_counter()
                for i in range(1, n_s):
                    model.track_stream_push(*hostf[i & 1])
                    model.track_stream_pop()
                model.track_stream_pop()
                dt_s = time.perf_counter() - t0
            out["config3_track"]["stream"] = {"value": (calls - 1) / dt_s, "unit": "fps", "frames": calls,
                                               "what": "fpb200_track_stream_push / pop, two frames in flight: double-buffered "
                                                       "upload on a copy stream, refined pose chained on the device"}
        except Exception as e:  # pragma: no cover
            out["config3_track"]["stream"] = {"unavailable": str(e)}
    except Exception as e:  # pragma: no cover
        out["config3_track"] = {"unavailable": str(e)}
    # ---- config 4 on one GPU: the 10k-triangle mesh (the N = 1 point of the sharded curve)
    if (args.workload or "mustard") != "mesh10k":
        try:
            mesh4, rgb4, depth4, mask4, gt4, _ = scene("mesh10k")
            m4 = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh4)], K, max_h=480, max_w=640,
                                               device=local_rank, max_hyps=N_HYP)
            s4 = torch.cuda.ExternalStream(m4.stream_ptr(), device=dev)
            r4, d4, k4 = _pin(rgb4), _pin(depth4), _pin(mask4)
            for _ in range(3):
                ok, _ = m4.Register(r4, d4, k4, mesh4.name, refine_itr=args.refine_itr)
                assert ok, m4.last_error
            m4.set_async(True)
            ms = timed_async_registers(m4, mesh4.name, args.steps, args.refine_itr, s4, torch)
            out["config4_mesh10k_1gpu"] = {"value": N_HYP * args.steps / (ms * 1e-3), "unit": UNIT,
                                           "ms_per_register": ms / args.steps, "triangles": int(mesh4.faces.shape[0])}
            m4.close()
        except Exception as e:  # pragma: no cover
            out["config4_mesh10k_1gpu"] = {"unavailable": str(e)}
    return out


def reference_cuda_preprocess(mesh, rgb, depth, mask, K, np):
    """Times the reference's own render path (oracle/_ref: vendored cudaraster + nvdiffrast + foundationpose_render.cu
    kernels, compiled in place; oracle/ref_render_shim.cu chains them as NvdiffrastRender + RenderProcess do) at N = 252
    on this GPU.  CV-CUDA (flip, warp) and TensorRT are absent, so this covers the render half of RenderAndTransform."""
    so = os.path.join(ROOT, "oracle", "_ref", "libref_kernels.so")
    if not os.path.exists(so):
        raise RuntimeError("oracle/_ref/libref_kernels.so not built")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(ROOT, "tests", "golden", "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    spec.loader.exec_module(mk)
    import fp_helpers as helpers
    from oracle import fp_oracle as O
    om = helpers.oracle_mesh(mesh)
    poses = O.sample_hypotheses(depth, mask, K).astype(np.float32)
    tfs, _ = O.crop_window_tf(poses, K, 1.2, om["diameter"])
    proj = O.projection_matrix(K, 480, 640)
    g = {"poses": poses, "mvp": np.stack([O._mat4_mul(proj, O._mat4_mul(O.GL_CAM_IN_CV_CAM, p)) for p in poses]).astype(np.float32),
         "bbox2d": O.bbox2d_from_tf(tfs).astype(np.float32), "vertices": om["vertices"], "normals": om["normals"],
         "texcoords": om["texcoords"], "faces": om["faces"], "texture": om["texture"], "diameter": np.float32(om["diameter"])}
    r = mk.run_render_sequence(g, repeat=10)
    return {"ms": r["ms"], "hypotheses": int(len(poses)), "kind": "reference",
            "what": "reference CUDA kernels in the reference's launch order: transform_points, generate_pose_clip, "
                    "cudaraster (4 kernels + host sync) + shader, interpolate x3, texture, transform_normals, refine_color, "
                    "clamp, threshold_and_downscale, concat (fp32 NHWC); the two cvcuda::Flip passes are not included"}


def micro_rooflines(model, torch, np, peak_tflops):
    """Live per-kernel rooflines of the tcgen05 kernels that are not conv_gemm: the fused encoder-layer tail and the
    self-attention core, each timed alone at N = 252 (CUDA events, 20 launches)."""
    from foundationpose_cpp_b200 import capi
    out = {}
    n = N_HYP
    rows = n * 400
    try:
        g = torch.Generator(device="cuda").manual_seed(0)
        att = (torch.randn(rows, 512, device="cuda", generator=g) * 0.5).half()
        x0 = torch.randn(rows, 512, device="cuda", generator=g).half()
        w = [(torch.randn(512, 512, device="cuda", generator=g) * 512 ** -0.5).half() for _ in range(3)]
        v = [torch.randn(512, device="cuda", generator=g) * 0.1 for _ in range(7)]
        partial, mean = torch.zeros(n * 25, 512, device="cuda"), torch.zeros(n, 512, device="cuda")
        d = capi.EncoderDesc()
        p = lambda t: C.c_void_p(t.data_ptr())
        d.att, d.x0, d.w_out, d.w1, d.w2 = p(att), p(x0), p(w[0]), p(w[1]), p(w[2])
        d.b_out, d.b1, d.b2, d.ln1_g, d.ln1_b, d.ln2_g, d.ln2_b = (p(t) for t in v)
        d.partial, d.mean_out, d.n_hyp, d.mode = p(partial), p(mean), n, 0
        ms = capi.debug_encoder_block(d, 20)
        fl = 3 * 2.0 * rows * 512 * 512
        out["encoder_block_kernel"] = {"ms": ms, "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_peak": fl / (ms * 1e-3) / 1e12 / peak_tflops,
                                       "flops": fl, "launches_per_register": 2,
                                       "what": "out-proj + LN1 + FFN + LN2 + token mean of one encoder layer, N = 252 (incl. the "
                                               "25-group mean kernel)"}
        del att, x0, partial, mean
        qkv = (torch.randn(rows, 1536, device="cuda", generator=g) * 0.5).half()
        o = torch.empty(rows, 512, device="cuda", dtype=torch.float16)
        ms = capi.debug_attention(qkv.data_ptr(), o.data_ptr(), n, 20)
        fl = 2.0 * n * 4 * 2 * 400 * 400 * 128
        out["attention_tc_kernel"] = {"ms": ms, "tflops": fl / (ms * 1e-3) / 1e12, "frac_of_peak": fl / (ms * 1e-3) / 1e12 / peak_tflops,
                                      "flops": fl, "launches_per_register": 3}
    except Exception as e:  # pragma: no cover
        out["unavailable"] = str(e)
    return out


def main():
    args = parse()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from foundationpose_cpp_b200 import api, sharding, weights

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    replicas = args.mode == "replicas" and world > 1
    sw, sr = (1, 0) if replicas else (world, rank)  # shard world / rank seen by the session
    kind = args.workload or ("mustard" if world == 1 else "mesh10k")
    mesh, rgb, depth, mask, gt, K = scene(kind, seed=rank if replicas else 0)
    wdir = tempfile.mkdtemp(prefix=f"fpb200_w{rank}_")
    rp, sp = weights.write_seeded(wdir)
    model = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh)], K, max_h=480, max_w=640,
                                          device=local_rank, max_hyps=N_HYP, shard_rank=sr, shard_world=sw)
    name = mesh.name
    # pinned host copies of the frame (what a capture pipeline would hand over)
    rgb_p, depth_p, mask_p = _pin(rgb), _pin(depth), _pin(mask)
    stream = torch.cuda.ExternalStream(model.stream_ptr(), device=dev)

    # ---- the exchange of a sharded Register: inside the library over NVLink peer memory
    exchange = "none (single GPU)" if sw == 1 else "peer memory"
    if sw > 1 and not args.host_gather:
        try:
            handles = sharding.gather_handles(model.exchange_export(), dist, dev)
            model.exchange_connect(handles)
        except Exception as e:
            exchange = f"host-driven NCCL all-gather (peer-memory exchange unavailable: {e})"
    elif sw > 1:
        exchange = "host-driven NCCL all-gather (--host-gather)"
    # all ranks must agree (a rank that fell back would never write its peers' flags)
    if sw > 1:
        flag = torch.tensor([1 if exchange == "peer memory" else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if int(flag.item()) == 0 and exchange == "peer memory":
            raise SystemExit("ranks disagree about the exchange path")
    p2p = exchange == "peer memory"
    cap = sharding.shard_capacity(N_HYP, sw)

    def gather():  # host-driven path only
        lp, nl, off, ap, na = model.feature_buffers()
        src = torch.as_tensor(_DevMem(lp, cap * 528), device=dev)
        dst = torch.as_tensor(_DevMem(ap, world * cap * 528), device=dev)
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(dst, src)

    profiling = [False]

    def step_resident():
        if (sw == 1 or p2p) and not profiling[0]:
            model.register_async(name, args.refine_itr)  # one CUDA-graph replay per Register
            return
        if sw == 1 or p2p:
            model.register_async(name, args.refine_itr)  # gemm-profile pass: the same launches, eager with event brackets
            return
        model.sample_hypotheses(None, name)
        for _ in range(args.refine_itr):
            model.refine_step()
        model.score_features()
        gather()
        model.score_finish_async()

    def step_e2e():
        if sw == 1 or p2p:
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
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- warm-up (also makes rgb/depth/mask resident); the clock sampler starts now so that it is already
    # streaming when the (possibly very short) timed region begins
    try:
        _uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        _uuid = None
    clocks = ClockSampler(local_rank, _uuid)
    model.set_async(sw > 1 and not p2p)
    warm = max(args.warmup, 3)
    for _ in range(warm):
        pose_w = step_e2e()
    model.set_async(True)
    for _ in range(3):  # eager, capture, replay of the resident-frame sequence
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
    model.set_async(sw > 1 and not p2p)
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
    whole_tflops = value * (f_ref * args.refine_itr + f_sco) / 1e12 / (world if not replicas else world)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak" if replicas else "strong",
        "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": WORKLOADS[kind], "hypotheses": N_HYP, "refine_itr": args.refine_itr,
                   "sharding": ("single GPU" if world == 1 else
                                f"{world} independent objects, one per GPU, no exchange" if replicas else
                                f"hypotheses split over {world} ranks; exchange of [N,528] f32 rows: {exchange}"),
                   "l2": "no explicit flush: each step streams ~2.6 GB of activations through HBM (>> 126 MB L2); "
                         "weights (66 MB) and the frame (2.4 MB) are legitimately hot",
                   "weights": "seeded random init of the published architecture (no checkpoint offline)",
                   "flops_per_hypothesis": {"refiner": f_ref, "scorer": f_sco},
                   "whole_job_tflops_per_gpu": whole_tflops,
                   "whole_job_frac_of_peak": whole_tflops / peak if peak else None,
                   "published_reference_rtx4060_hyp_s": 705.6},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(rgb.nbytes + depth.nbytes + mask.nbytes),
                "d2h_bytes_per_step": 72, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel + stem_conv_kernel (tcgen05 implicit GEMM family: the 15 conv layers + the QKV projections)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     "traffic": traffic, "peak_source": peak_src, "launches_timed": int(gemm_launches),
                     "kernel_ms_per_step": gemm_ms / args.steps, "kernel_share_of_step": gemm_ms / prof_ms if prof_ms else None,
                     "timing": "second pass of the same K steps with CUDA events around every conv_gemm launch"},
    }
    # ---- exchange time of the sharded Register, alone (all ranks in lockstep)
    if p2p:
        model.set_async(True)
        barrier()
        ex_ms = max_over_ranks(model.time_stage(2, 50))
        barrier()
        out["collective_ms"] = ex_ms
        out["config"]["exchange"] = ("peer-memory exchange inside the library: every rank stores its [n_local,528] f32 rows "
                                     "into every rank's region over NVLink + release flags; collective_ms = begin + push + "
                                     "collect kernels timed alone, max over ranks")
    # ---- sub-records of the other BASELINE configs
    if not args.no_extras:
        configs = {}
        if world == 1:
            configs.update(extra_single_gpu(args, model, mesh, rgb_p, depth_p, mask_p, gt, K, rp, sp, local_rank, torch, np))
            out["roofline"]["other_kernels"] = micro_rooflines(model, torch, np, peak)
        elif not replicas:
            # config 5: one object per GPU (8 meshes x 252 hypotheses at 8 GPUs), no exchange at all
            try:
                mesh5, rgb5, depth5, mask5, gt5, _ = scene("mesh10k", seed=rank)
                m5 = api.CreateFoundationPoseModel(rp, sp, [api.ArrayMeshLoader.from_mesh_data(mesh5)], K, max_h=480,
                                                   max_w=640, device=local_rank, max_hyps=N_HYP)
                s5 = torch.cuda.ExternalStream(m5.stream_ptr(), device=dev)
                r5, d5, k5 = _pin(rgb5), _pin(depth5), _pin(mask5)
                for _ in range(3):
                    ok, _ = m5.Register(r5, d5, k5, mesh5.name, refine_itr=args.refine_itr)
                    assert ok, m5.last_error
                m5.set_async(True)
                for _ in range(2):
                    m5.register_async(mesh5.name, args.refine_itr)
                m5.wait()
                barrier()
                ms5 = max_over_ranks(timed_async_registers(m5, mesh5.name, args.steps, args.refine_itr, s5, torch))
                barrier()
                configs["config5_one_object_per_gpu"] = {"value": world * N_HYP * args.steps / (ms5 * 1e-3), "unit": UNIT,
                                                         "ms_per_register": ms5 / args.steps, "objects": world,
                                                         "scaling": "weak", "exchange": "none"}
                m5.close()
            except Exception as e:  # pragma: no cover
                configs["config5_one_object_per_gpu"] = {"unavailable": str(e)}
        out["configs"] = configs
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import torch as _t
        _t.set_num_threads(os.cpu_count() or 1)
        ctx = cpu_context(kind)
        cpu_reference_step(ctx, 1, args.refine_itr)
        n_s, t_s, t_h, t0 = 0, 0.0, 0.0, time.perf_counter()
        while time.perf_counter() - t0 < 12.0:
            a, b = cpu_reference_step(ctx, 2, args.refine_itr)
            t_s += a
            t_h += b
            n_s += 2
        calls = n_s // 2
        full_step_s = t_s / calls + N_HYP * (t_h / n_s)
        out["cpu_baseline"] = {"value": N_HYP / full_step_s, "unit": UNIT, "cores": ctx["cores"], "kind": "port",
                               "sample": f"{calls} calls of (whole sampler + 2 of the 252 hypotheses through refine / score / "
                                         f"argmax) in {time.perf_counter() - t0:.1f} s; value = 252 / (sampler + 252 x "
                                         f"per-hypothesis time)",
                               "host_pre_post": host_prepost_baseline(ctx)}
    elif rank == 0:
        out["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(out))
    model.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
