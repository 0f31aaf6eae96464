#!/usr/bin/env python3
"""bench.py -- frames/s of the per-frame pixel hot path on N B200s, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1], "C2"): ZED 2208x1242 stereo frames through balance() (default
flags, modules/color_balance.py:93-96) and BGR2LAB (utils/color.py:26).  One step = one batch of
16 frames (8 stereo pairs) per GPU taken from a ring of 64 distinct synthetic frames (526 MB, > L2;
consecutive steps use different batches).  Frames shard by index across GPUs with no collective
("weak" scaling: 16 frames per GPU per step).

One JSON line on stdout (rank 0).  `value`: device-resident frames/s, CUDA events on the library's
stream.  `e2e`: same stage through bv_stage_host with pinned HOST buffers, H2D + D2H inside the
timed region.  `roofline`: dominant kernel, per-launch CUDA-event time from the library's own
profiler; `stage_roofline`: the whole step against the algorithmic 6 B/px.  `cpu_baseline`: the
reference's compiled process_frame + cv2.cvtColor on the host cores (reported, not a target).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W = 1242, 2208
BATCH = 16
RING = 64
BPP_C2 = 6          # BGR in (3) + LAB image out (3)
METRIC = "frames/sec at 2208x1242"


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self, wait_s=5.0):
        """Starts nvidia-smi and waits for its first sample, so that even a short timed region is
        covered (nvidia-smi needs ~0.5 s to come up)."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            t0 = time.time()
            while not self.lines and time.time() - t0 < wait_s:
                time.sleep(0.02)
        except Exception:  # noqa: BLE001
            self.proc = None

    def mark(self):
        """Index of the next sample: samples[mark_a:mark_b] belong to a timed region."""
        return len(self.lines)

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self, first=0, last=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # the device-resident region can be shorter than one sampling period: widen by one sample
        # on each side, the rest of the window (profile + end-to-end legs) is under load as well
        lines = self.lines[max(0, first - 1):(None if last is None else last + 1)] or self.lines
        for ln in lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_ring(n, h=H, w=W, seed0=2000):
    from oracle import synth  # synthetic inputs only (generators live with the test infrastructure)
    return np.stack([synth.gen_underwater(h, w, seed0 + i) for i in range(n)])


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_step(frames, pool):
    import cv2
    from oracle import ref_balance, color_balance_np

    def one(img):
        bal = ref_balance.balance(img) if ref_balance.available() else color_balance_np.process_frame_np(img)
        return cv2.cvtColor(bal, cv2.COLOR_BGR2LAB)
    return list(pool.map(one, frames))


def cpu_baseline(frames, cores):
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_balance
    cv2.setNumThreads(1)
    with ThreadPoolExecutor(max_workers=cores) as pool:
        cpu_reference_step(frames[:min(len(frames), cores)], pool)     # warm-up
        t0 = time.perf_counter()
        cpu_reference_step(frames, pool)
        dt = time.perf_counter() - t0
    cv2.setNumThreads(0)
    return {"value": len(frames) / dt, "unit": "frames/s", "cores": cores,
            "kind": "reference" if ref_balance.available() else "port",
            "sample": "%d frames 2208x1242: compiled reference process_frame (default flags, marshalled as "
                      "modules/color_balance.py:93-110) + cv2.cvtColor(BGR2LAB), one frame per thread, %.2f s wall"
                      % (len(frames), dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ref_balance
    cores = os.cpu_count() or 1
    per_step = max(2, min(BATCH, cores))
    ring = make_ring(per_step * 2)
    cv2.setNumThreads(1)
    times = []
    with ThreadPoolExecutor(max_workers=cores) as pool:
        for s in range(args.warmup + args.steps):
            frames = ring[(s % 2) * per_step:(s % 2 + 1) * per_step]
            t0 = time.perf_counter()
            cpu_reference_step(frames, pool)
            dt = time.perf_counter() - t0
            if s >= args.warmup:
                times.append(dt)
    total = sum(times)
    value = per_step * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "C2: ZED 2208x1242 stereo frames, balance() default flags -> BGR2LAB image "
                               "(BASELINE.json configs[1])",
                   "reference_sample": "each step = %d frames on the host cores (compiled reference process_frame + "
                                       "cv2.cvtColor), one frame per thread" % per_step,
                   "frames_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores,
                         "kind": "reference" if ref_balance.available() else "port",
                         "sample": "%d steps x %d frames, one frame per thread over %d threads" % (args.steps, per_step, cores)},
        "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def time_device(ctx, fn, steps, warmup):
    """fn(step) enqueues one step on ctx's stream.  CUDA events on that stream."""
    import torch
    for s in range(warmup):
        fn(s)
    ctx.sync()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    for s in range(steps):
        fn(warmup + s)
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    ctx.sync()
    return e0.elapsed_time(e1) / 1e3


def side_workloads(ctx, peak_gbs):
    """Short device-resident measurements of the other BASELINE configs (not the headline)."""
    import torch
    from oracle import synth
    res = {}

    def fps(name, fn, frames, bpp_px, steps=10, warmup=3):
        dt = time_device(ctx, fn, steps, warmup)
        v = frames * steps / dt
        res[name] = {"frames_per_s": v, "ms_per_step": 1e3 * dt / steps,
                     "algorithmic_gbs": bpp_px * v / 1e9, "frac_of_hbm": bpp_px * v / 1e9 / peak_gbs}
    # north-star stage at 2208x1242: balance -> HSV -> inRange -> OPEN 5x5, mask out (4 B/px)
    ring = ctx.upload(np.stack([synth.gen_underwater(H, W, 3000 + i) for i in range(16)]))
    desc = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(0, 40, 60), hi=(179, 255, 255), morph=[("open", 5, 5, 1)])
    out = {}
    fps("fused_balance_hsv_inrange_open_2208x1242", lambda s: out.update(ctx.stage(desc, ring, want=("mask",), out=out)),
        16, 4 * H * W)
    # C3: 1920x1080 HSV inRange -> OPEN -> CCL + moments (8 B/px)
    ring3 = ctx.upload(np.stack([synth.gen_underwater(1080, 1920, 3100 + i) for i in range(16)]))
    d3 = ctx.make_stage(cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
    o3 = {}
    fps("c3_hsv_inrange_open_label_1920x1080",
        lambda s: o3.update(ctx.stage(d3, ring3, want=("mask", "labels", "blobs"), max_blobs=4096, out=o3)), 16, 8 * 1080 * 1920)
    # C5: 3840x2160 balance -> HSV -> inRange -> OPEN -> label (8 B/px), 8 streams
    ring5 = ctx.upload(np.stack([synth.gen_underwater(2160, 3840, 3200 + i) for i in range(8)]))
    d5 = ctx.make_stage(balance={}, cvt="bgr2hsv", lo=(10, 20, 60), hi=(30, 100, 255), morph=[("open", 5, 5, 1)], label=True)
    o5 = {}
    fps("c5_balance_threshold_label_3840x2160",
        lambda s: o5.update(ctx.stage(d5, ring5, want=("mask", "labels", "blobs"), max_blobs=8192, out=o5)), 8, 8 * 2160 * 3840)
    # C4: 16 frames -> letterbox 640x640 fp16
    imgs = [ctx.upload(synth.gen_underwater(H, W, 3300 + i)) for i in range(16)]
    bytes_batch = 16 * H * W * 3 + 16 * 3 * 640 * 640 * 2
    dt = time_device(ctx, lambda s: ctx.letterbox(imgs), 10, 3)
    res["c4_letterbox_16x2208x1242_to_640_fp16"] = {"images_per_s": 160 / dt, "ms_per_step": 1e2 * dt,
                                                   "algorithmic_gbs": bytes_batch * 10 / dt / 1e9,
                                                   "frac_of_hbm": bytes_batch * 10 / dt / 1e9 / peak_gbs}
    # pure streaming kernels (what the HBM roofline looks like on this path when the arithmetic is light)
    rgba = ctx.upload(np.random.default_rng(0).integers(0, 256, (16, H, W, 4), dtype=np.uint8))
    gray_out = {}
    dt = time_device(ctx, lambda s: ctx.rgba_to_rgb(rgba), 10, 3)
    res["stream_rgba_to_rgb_16x2208x1242"] = {"frames_per_s": 160 / dt, "ms_per_step": 1e2 * dt,
                                              "algorithmic_gbs": 7 * H * W * 160 / dt / 1e9,
                                              "frac_of_hbm": 7 * H * W * 160 / dt / 1e9 / peak_gbs}
    dt = time_device(ctx, lambda s: ctx.cvt_color(ring, "bgr2gray"), 10, 3)
    res["stream_bgr2gray_16x2208x1242"] = {"frames_per_s": 160 / dt, "ms_per_step": 1e2 * dt,
                                           "algorithmic_gbs": 4 * H * W * 160 / dt / 1e9,
                                           "frac_of_hbm": 4 * H * W * 160 / dt / 1e9 / peak_gbs}
    del ring, ring3, ring5, imgs, rgba
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cuauv_vision_pipeline_b200 as bv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa_cpus = []
    if world > 1:
        torch.cuda.set_device(local)
        # one process per GPU: stay on the CPUs next to that GPU so the pinned buffers of the end-to-end leg
        # are allocated on its NUMA node (at N=1 the process keeps every core for the CPU baseline leg)
        from cuauv_vision_pipeline_b200.sharding import bind_to_gpu_numa
        numa_cpus = bind_to_gpu_numa(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = bv.Context(local)
    peak_gbs, peak_src = read_peaks()

    # synthetic ring, distinct per rank (frame f of the global stream goes to rank f mod world)
    ring_np = make_ring(RING, seed0=2000 + 100 * rank)
    ring = ctx.upload(ring_np)
    n_batches = RING // BATCH
    desc = ctx.make_stage(balance={}, cvt="bgr2lab")
    out = {}

    def step(s):
        b = s % n_batches
        out.update(ctx.stage(desc, ring[b * BATCH:(b + 1) * BATCH], want=("converted",), out=out))

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ----
    for s in range(args.warmup):
        step(s)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        for s in range(args.warmup):   # keep the GPU busy while the sampler came up
            step(s)
        ctx.sync()
    barrier()
    mark_a = sampler.mark()
    launches0 = ctx.launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(ctx.torch_stream):
        e0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    with torch.cuda.stream(ctx.torch_stream):
        e1.record()
    barrier()
    elapsed = e0.elapsed_time(e1) / 1e3
    launches = ctx.launches - launches0
    if world > 1:
        t = torch.tensor([elapsed], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())
    value = world * BATCH * args.steps / elapsed

    # ---- per-kernel timing (library profiler: CUDA events around every launch) ----
    ctx.profile(True)
    for s in range(4):
        step(s)
    prof = ctx.profile_dump()
    ctx.profile(False)
    total_ms = sum(v["ms"] for v in prof.values()) or 1.0
    dom = max(prof, key=lambda k: prof[k]["ms"])
    # one launch of any pass covers one L2-sized chunk of frames
    chunk_frames = max(1, int(os.environ.get("BV_L2_CHUNK_MB", "33")) * (1 << 20) // (H * W * 3))
    chunk_frames = min(chunk_frames, BATCH)
    dom_ms = prof[dom]["ms"] / prof[dom]["launches"]
    achieved = BPP_C2 * H * W * chunk_frames / (dom_ms / 1e3) / 1e9
    traffic, traffic_src = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "r01_traffic.json")) as f:
            t = json.load(f)
        if dom in t["kernels"]:
            per_frame = t["kernels"][dom]["dram_bytes_per_launch"] / t["kernels"][dom]["frames_per_launch"]
            traffic, traffic_src = per_frame * chunk_frames, t["source"]
    except Exception:  # noqa: BLE001
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "kernel_ms_per_launch": dom_ms, "kernel_share_of_step": prof[dom]["ms"] / total_ms,
                "algorithmic_bytes_per_launch": BPP_C2 * H * W * chunk_frames,
                "note": "achieved = 6 B/px (BGR in + LAB out) x frames in one launch / that kernel's CUDA-event time"}
    stage_gbs = BPP_C2 * H * W * value / world / 1e9
    stage_roofline = {"achieved": stage_gbs, "peak": peak_gbs, "unit": "GB/s", "frac": stage_gbs / peak_gbs,
                      "per_kernel_ms": {k: round(v["ms"] / 4, 4) for k, v in sorted(prof.items())}}

    # ---- end to end through the host-buffer entry point (pinned memory) ----
    pin_in = bv.PinnedArray((2, BATCH, H, W, 3))
    pin_in.array[0] = ring_np[:BATCH]
    pin_in.array[1] = ring_np[BATCH:2 * BATCH]
    pin_out = bv.PinnedArray((BATCH, H, W, 3))
    host_out = {"converted": pin_out.array}
    e2e_steps = max(3, min(args.steps, 10))
    for s in range(2):
        ctx.stage_host(desc, pin_in.array[s % 2], want=("converted",), out=host_out)
    barrier()
    t0 = time.perf_counter()
    for s in range(e2e_steps):
        ctx.stage_host(desc, pin_in.array[s % 2], want=("converted",), out=host_out)
    barrier()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    # single-frame latency of the drop-in call (what a module's process() pays per frame)
    one_out = {"converted": pin_out.array[:1]}
    for s in range(3):
        ctx.stage_host(desc, pin_in.array[0][:1], want=("converted",), out=one_out)
    t0 = time.perf_counter()
    for s in range(20):
        ctx.stage_host(desc, pin_in.array[0][s % BATCH:s % BATCH + 1], want=("converted",), out=one_out)
    single_ms = (time.perf_counter() - t0) / 20 * 1e3
    # the box's own host<->device copy ceiling with both directions busy, measured with the same pinned buffers
    # (it differs between boxes of the pool: 62-99 GB/s seen), so that the end-to-end number can be read against it
    pcie = None
    if world == 1:
        d_in = torch.empty(pin_in.array[0].shape, dtype=torch.uint8, device="cuda")
        d_out = torch.empty(pin_out.array.shape, dtype=torch.uint8, device="cuda")
        t_in, t_out = torch.from_numpy(pin_in.array[0]), torch.from_numpy(pin_out.array)
        s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
        torch.cuda.synchronize()

        def both_ways():
            with torch.cuda.stream(s_up):
                d_in.copy_(t_in, non_blocking=True)
            with torch.cuda.stream(s_down):
                t_out.copy_(d_out, non_blocking=True)
        both_ways()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            both_ways()
        torch.cuda.synchronize()
        both_dt = (time.perf_counter() - t0) / 5
        pcie = {"both_directions_gbs": (t_in.numel() + t_out.numel()) / both_dt / 1e9,
                "ceiling_frames_per_s": BATCH / both_dt}
        del d_in, d_out
    clocks = sampler.stop(mark_a, None) if rank == 0 else None
    if clocks is not None:
        clocks["window"] = "device-resident timed region + per-kernel profile + end-to-end leg"
    e2e = {"value": world * BATCH * e2e_steps / e2e_dt, "unit": "frames/s",
           "h2d_bytes_per_step": BATCH * H * W * 3, "d2h_bytes_per_step": BATCH * H * W * 3,
           "api": "bv_stage_host (C ABI, pinned host buffers, blocking)", "steps": e2e_steps,
           "single_frame_latency_ms": single_ms,
           "pcie_note": "8.23 MB in + 8.23 MB out per frame, copied in both directions at once; `pcie` is this box's own "
                        "ceiling for that (plain cudaMemcpyAsync of the same pinned buffers, no kernels)"}
    if pcie is not None:
        pcie["e2e_frac_of_ceiling"] = e2e["value"] / pcie["ceiling_frames_per_s"]
        e2e["pcie"] = pcie

    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = cpu_baseline(list(ring_np[:min(32, max(8, cores))]), cores) if (world == 1 and not args.no_cpu) else None
        others = side_workloads(ctx, peak_gbs) if (world == 1 and not args.no_side) else None
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C2: ZED 2208x1242 stereo frames, balance() default flags -> BGR2LAB image "
                                   "(BASELINE.json configs[1])",
                       "frames_per_step_per_gpu": BATCH, "ring_frames": RING,
                       "l2": "inputs larger than L2 (131 MB per step from a 526 MB ring; consecutive steps use "
                             "different batches)",
                       "parallelism": "frames sharded by index over %d GPU(s), no collective" % world,
                       "host_affinity": ("%d CPUs local to each GPU" % len(numa_cpus)) if numa_cpus else "default"},
            "roofline": roofline, "stage_roofline": stage_roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "other_workloads": others,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-side", action="store_true", help="skip the short measurements of the other configs")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (tuning sweeps only)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
