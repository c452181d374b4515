#!/usr/bin/env python
"""bench.py -- proposals/s of WSOVOD's region-scoring hot path on B200 (one process per GPU).

    python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N   # N ranks, weak scaling
    python bench.py --impl reference --steps 3 --warmup 1          # the reference's CPU path (host cores)

Workload (BASELINE.json configs[1], "c2"): COCO WSOVOD_WSR_18_DC5 inference, 8 images / GPU of
688x1024 (res5 map 512x86x128, stride 8), 4000 proposals / image, 80 concepts + background, D=768,
T=50, score_thresh 1e-5, nms 0.3, 100 detections / image.  Synthetic, seeded (wsovod_b200/synth.py).

One step = ROI pool (+objectness scale) -> [box-head FCs: out of scope, embeddings are synthetic]
-> region x concept alignment + softmax (tcgen05 TF32) -> per-class NMS + top-100  for the GPU's 8
images.  `value`: inputs resident in HBM.  `e2e`: the same step through the C-ABI host entry point
(wsovod_b200_infer_host) from pinned HOST buffers, H2D/D2H inside the timed region.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "proposals/sec of region-scoring path"
UNIT = "proposals/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, source="fallback (B200_PROFILING.md)")


def load_traffic(with_arg):
    """dram__bytes_read.sum + dram__bytes_write.sum of the pooling kernel, per launch, from the committed
    `ncu --set full` capture of this workload (profiles/roofline_traffic.json); None if absent."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if not os.path.exists(p):
        return None
    d = json.load(open(p))
    return d.get("roi_pool+argmax" if with_arg else "roi_pool")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        t0, t1 = getattr(self, "t0", 0.0), getattr(self, "t1", float("inf"))
        inside = [r for t, r in self.rows if t0 <= t <= t1 + 0.05]
        # a short timed region can fall between two nvidia-smi samples: then use every sample taken while
        # the GPU was running this benchmark's steps (warm-up + timed region + per-kernel timing)
        rows = inside if len(inside) >= 3 else [r for _, r in self.rows]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except Exception:  # noqa: BLE001
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    samples=len(sm), samples_in_timed_region=len(inside), reasons=sorted(reasons))


def cpu_reference_leg(w, budget_s=20.0):
    """The reference's CPU path (oracle/cpu_path.py) on a bounded sample of the same workload."""
    from oracle import cpu_path
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # size the sample from a tiny probe so that one pass costs ~10-30 s at most
    dt, n, *_ = cpu_path.run_slice(w, images=1, proposals=100)
    per_prop = dt / n
    props = int(max(200, min(w["R"], budget_s / max(per_prop, 1e-9))))
    images = 1
    if props >= w["R"]:
        props = w["R"]
        images = int(max(1, min(w["N"], budget_s / max(per_prop * props, 1e-9))))
    return cpu_path, cores, images, props


def run_reference(args, w, rank):
    if rank != 0:
        return
    # the whole --steps K run has to end within a few minutes: ~150 s of CPU work in total
    cpu_path, cores, images, props = cpu_reference_leg(w, budget_s=max(2.0, min(20.0, 150.0 / max(args.steps, 1))))
    for _ in range(args.warmup):
        cpu_path.run_slice(w, images=1, proposals=min(props, 200))
    t, n = 0.0, 0
    for _ in range(args.steps):
        dt, k, *_ = cpu_path.run_slice(w, images=images, proposals=props)
        t += dt
        n += k
    v = n / t
    sample = f"{images} image(s) x {props} proposals of c2 per step (pool+scale, align+softmax, NMS+top100)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "c2: COCO WSR_18_DC5 inference slice, 8 img x 4000 proposals/GPU, K=80, D=768",
                   "device": "cpu"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--pool-argmax", type=int, default=0,
                    help="1: the pooling kernel also emits argmax (training with a trainable backbone); the "
                         "reference's frozen-backbone inference never reads it (SURVEY fact 6)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    from wsovod_b200 import shard, synth
    rank, local_rank, world = shard.env_world()
    w = synth.workload(args.config, seed=1234, rank=rank)
    if args.impl == "reference":
        run_reference(args, w, rank)
        return

    from wsovod_b200 import _lib, ops
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device: wsovod_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly one JSON line: NCCL writes its "NCCL version ..." banner to fd 1 when the first
    # communicator is created, so the init and the first collective run with fd 1 pointed at stderr
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        rank, local_rank, world = shard.init("nccl")
        shard.barrier()
        torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    N, C, H, W, R, K, D = (w[k] for k in "NCHWRKD")
    M = N * R
    feat, rois, obj = w["features"].to(dev), w["rois"].to(dev), w["objectness"].to(dev)
    emb, text = w["region_emb"].to(dev), w["text_emb"].to(dev)
    off = torch.tensor(w["offsets"], dtype=torch.int64, device=dev)
    sizes = w["image_sizes"].to(dev)
    boxes = rois[:, 1:].contiguous()
    with_arg = bool(args.pool_argmax)

    def step():
        pooled, _ = ops.roi_pool(feat, rois, w["spatial_scale"], 7, obj, 1.0, with_arg)
        _, probs = ops.align(emb, text, w["temperature"], 1, True, None, ops.ALIGN_TF32, False, True)
        det = ops.detections(probs, boxes, off, sizes, R, w["score_thresh"], w["nms_thresh"], w["topk"],
                             ops.IOU_TV_CUDA)
        return pooled, det

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        out = step()
    torch.cuda.synchronize()
    # ---- timed region: exactly `steps` steps, device time, barrier + synchronize on both sides ------
    shard.barrier()
    torch.cuda.synchronize()
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    e0.record()
    for _ in range(args.steps):
        out = step()
    e1.record()
    torch.cuda.synchronize()
    sampler.window(t_wall0, time.time())
    shard.barrier()
    ms_total = shard.max_over_ranks(e0.elapsed_time(e1), dev)
    launches = _lib.launch_count() - launches0
    value = world * M * args.steps / (ms_total * 1e-3)

    # ---- per-kernel device times (CUDA events on the launching stream), same inputs ----------------
    def ktime(fn, iters):
        """device time of one call of `fn`: `iters` calls captured in a CUDA graph and replayed, so that the
        sub-100-us ops are not timed at the pace of their Python wrappers; plain loop if capture fails"""
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(iters):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            best = None
            for _ in range(3):
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                t = a.elapsed_time(b) / iters
                best = t if best is None else min(best, t)
            del g
            return best
        except Exception:  # noqa: BLE001
            torch.cuda.synchronize()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    it = max(args.steps // 2, 3)
    probs = ops.align(emb, text, w["temperature"], 1, True, None, ops.ALIGN_TF32, False, True)[1]
    t_pool = ktime(lambda: ops.roi_pool(feat, rois, w["spatial_scale"], 7, obj, 1.0, with_arg), it)
    t_pool_arg = ktime(lambda: ops.roi_pool(feat, rois, w["spatial_scale"], 7, obj, 1.0, True), it)
    t_align = ktime(lambda: ops.align(emb, text, w["temperature"], 1, True, None, ops.ALIGN_TF32, False, True), it)
    t_det = ktime(lambda: ops.detections(probs, boxes, off, sizes, R, w["score_thresh"], w["nms_thresh"], w["topk"],
                                         ops.IOU_TV_CUDA), it)
    clocks = sampler.stop() if rank == 0 else None
    peaks = load_peaks()
    out_bytes = M * C * 49 * 4
    pool_bytes = out_bytes * (2 if with_arg else 1) + feat.numel() * 4 + M * 20      # DESIGN.md "Kernel 1"
    pool_gbs = pool_bytes / (t_pool * 1e-3) / 1e9
    kernels = {
        "roi_pool": {"ms": t_pool, "algorithmic_bytes": pool_bytes, "GBps": pool_gbs, "frac_hbm": pool_gbs / peaks["hbm_gbs"]},
        "roi_pool+argmax": {"ms": t_pool_arg, "algorithmic_bytes": 2 * out_bytes + feat.numel() * 4 + M * 20,
                            "GBps": (2 * out_bytes + feat.numel() * 4 + M * 20) / (t_pool_arg * 1e-3) / 1e9},
        "align_tf32+softmax": {"ms": t_align, "algorithmic_bytes": M * D * 4 + M * (K + 1) * 4 + K * D * 4,
                               "GBps": (M * D * 4 + M * (K + 1) * 4 + K * D * 4) / (t_align * 1e-3) / 1e9,
                               "TFLOPs": 2.0 * M * D * (K + 1) / (t_align * 1e-3) / 1e12},
        "nms+top100": {"ms": t_det, "candidates": int((probs[:, :-1] > w["score_thresh"]).sum()),
                       "algorithmic_bytes": M * (K + 1) * 4 + M * 16},
    }
    for k in ("roi_pool+argmax", "align_tf32+softmax"):
        kernels[k]["frac_hbm"] = kernels[k]["GBps"] / peaks["hbm_gbs"]

    # ---- e2e: host buffers through the C-ABI entry point --------------------------------------------
    L = _lib.lib()
    pin = lambda t: t.contiguous().pin_memory()  # noqa: E731
    h_feat, h_rois, h_obj = pin(w["features"]), pin(w["rois"]), pin(w["objectness"])
    h_emb, h_text, h_sizes = pin(w["region_emb"]), pin(w["text_emb"]), pin(w["image_sizes"])
    h_off = pin(torch.tensor(w["offsets"], dtype=torch.int64))
    topk = w["topk"]
    h_db = torch.empty(N, topk, 4).pin_memory()
    h_ds = torch.empty(N, topk).pin_memory()
    h_dc = torch.empty(N, topk, dtype=torch.int64).pin_memory()
    h_dr = torch.empty(N, topk, dtype=torch.int64).pin_memory()
    h_cnt = torch.empty(N, dtype=torch.int64).pin_memory()
    arena_bytes = L.wsovod_b200_infer_host_arena(N, C, H, W, M, D, K, 7, topk, int(with_arg))
    arena = torch.empty(arena_bytes, dtype=torch.uint8, device=dev)
    P = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    copy_s = torch.cuda.Stream(dev)
    copy_stream = ctypes.c_void_p(copy_s.cuda_stream)

    def e2e_step():
        rc = L.wsovod_b200_infer_host(P(h_feat), N, C, H, W, P(h_rois), P(h_obj), M, P(h_off), P(h_sizes), P(h_emb),
                                      P(h_text), D, K, w["spatial_scale"], 7, w["temperature"], w["score_thresh"],
                                      w["nms_thresh"], topk, 1, 1, int(with_arg), P(h_db), P(h_ds), P(h_dc), P(h_dr),
                                      P(h_cnt), P(arena), arena_bytes, None, stream, copy_stream)
        _lib.check(rc, "infer_host")
        torch.cuda.current_stream(dev).synchronize()      # the host reads the step's detections
        return int(h_cnt.sum())

    for _ in range(3):
        e2e_step()
    shard.barrier()
    torch.cuda.synchronize()
    e2e_steps = max(args.steps // 2, 3)
    t0 = time.perf_counter()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(e2e_steps):
        ndet = e2e_step()
    b.record()
    torch.cuda.synchronize()
    e2e_ms = shard.max_over_ranks(a.elapsed_time(b), dev)
    e2e_wall = time.perf_counter() - t0
    e2e_value = world * M * e2e_steps / (e2e_ms * 1e-3)
    h2d = sum(t.numel() * t.element_size() for t in (h_feat, h_rois, h_obj, h_emb, h_text, h_sizes, h_off))
    d2h = sum(t.numel() * t.element_size() for t in (h_db, h_ds, h_dc, h_dr, h_cnt))
    # device-resident and host-buffer paths must agree on the detections
    same = bool(torch.equal(out[1]["det_scores"].cpu(), h_ds) and torch.equal(out[1]["det_rows"].cpu(), h_dr))

    cpu_base = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_path, cores, images, props = cpu_reference_leg(w, budget_s=12.0)
        dt, n, *_ = cpu_path.run_slice(w, images=images, proposals=props)
        cpu_base = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                    "sample": f"{images} image(s) x {props} proposals of c2, one pass: torchvision CPU roi_pool + "
                              f"objectness scale, ATen normalize/mm/softmax, torchvision CPU batched_nms + top100"}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (pool, softmax, NMS) / tf32 (alignment contraction)", "data": "synthetic",
            "config": {"workload": "c2: COCO WSR_18_DC5 inference slice, 8 img x 4000 proposals/GPU, K=80, D=768",
                       "global_proposals_per_step": world * M, "pool_argmax": with_arg,
                       "l2": "inputs+outputs per step (3.5 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": f"dp{world} (images sharded, no data-path collective)"},
            "roofline": {"bound": "hbm", "kernel": "roi_pool7_pyr_kernel (ROI max-pool, block-max planes)", "achieved": pool_gbs,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": pool_gbs / peaks["hbm_gbs"],
                         "peak_source": peaks["source"], "traffic": load_traffic(with_arg),
                         "share_of_step": t_pool / (ms_total / args.steps)},
            "kernels": kernels,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms / e2e_steps, "wall_ms_per_step": 1e3 * e2e_wall / e2e_steps,
                    "matches_device_path": same, "detections_last_step": ndet},
            "gpu_launches": int(launches),
            "clocks": clocks,
        }
        if cpu_base:
            line["cpu_baseline"] = cpu_base
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
