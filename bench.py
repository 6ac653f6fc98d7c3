#!/usr/bin/env python
"""bench.py — front-end throughput of the B200 ORB path on BASELINE.json's KITTI-shaped stereo workload.

One step = one pass of the hot path over one batch of synthetic stereo pairs: detect (pyramid, FAST
cells, quadtree, orientation, blur, rBRIEF) on every image of the batch + left-right windowed Hamming
matching of every pair.  `value` is measured with the frames already resident in HBM (CUDA events on
the library's own stream), `e2e` through the C-ABI calls with pinned HOST buffers (H2D of the frames and
D2H of keypoints, descriptors and matches inside the timed region).  Sequences shard across GPUs with no
collective (weak scaling: every rank processes its own batch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl b200|reference]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NFEAT, NLEVELS = 376, 1241, 2000, 8
PITCH = 1280                      # 16-byte aligned row pitch of the device-resident frames
L2_BYTES = 126e6
NCU_DRAM_BYTES_PER_FRAME = 870.8e6 / 256     # pyramid + FAST, measured (see roofline.traffic)


def level_pixels(h, w, nlevels=8):
    s = np.float32(1.0); tot = 0
    for _ in range(nlevels):
        inv = np.float32(1.0) / s
        tot += int(np.rint(np.float32(w) * inv)) * int(np.rint(np.float32(h) * inv))
        s = np.float32(float(s) * float(np.float32(1.2)))
    return tot


def make_pairs(n_pairs, seed, n_scenes=8):
    """n_pairs stereo pairs [2*n_pairs, H, W]: left = crop of a scene, right = same crop shifted by dx."""
    from nav24_b200.synth import canvas, frame_from_canvas
    rng = np.random.default_rng(seed)
    scenes = [canvas(H, W, seed * 100 + s) for s in range(n_scenes)]
    out = np.empty((2 * n_pairs, H, W), np.uint8)
    for p in range(n_pairs):
        c = scenes[p % n_scenes]
        sx = int(rng.integers(-60, 0)); sy = int(rng.integers(-60, 60)); dx = int(rng.integers(4, 64))
        out[2 * p] = frame_from_canvas(c, H, W, (sx + dx, sy), noise_seed=seed * 7919 + 2 * p)
        out[2 * p + 1] = frame_from_canvas(c, H, W, (sx, sy), noise_seed=seed * 7919 + 2 * p + 1)
    return out


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill(); out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_frontend_threads(frames, n_threads, pairs_per_thread, nfeat=NFEAT):
    """Runs the CPU oracle front end (detect L, detect R, windowed match) on n_threads host threads.
    Returns (frames processed, keypoints, wall seconds). The ctypes calls release the GIL."""
    from oracle import orb_oracle as oo
    oo.lib()
    grid = oo.grid_for(W, H)
    n_pairs_avail = len(frames) // 2
    stats = [None] * n_threads

    def work(t):
        o = oo.OrbOracle(nfeat)
        nk = 0
        for i in range(pairs_per_thread):
            p = (t * pairs_per_thread + i) % n_pairs_avail
            _, k1, d1 = o.detect(frames[2 * p]); _, k2, d2 = o.detect(frames[2 * p + 1])
            ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
            oo.match_window(k1, ud1, d1, k2, ud2, d2, grid)
            nk += len(k1) + len(k2)
        stats[t] = nk

    ths = [threading.Thread(target=work, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    return 2 * pairs_per_thread * n_threads, sum(stats), dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's host threads to the cores of the NUMA node its GPU hangs off, BEFORE any pinned buffer is
    allocated (first touch places the pages there): with 8 ranks per box the host->device copies otherwise cross the
    socket interconnect.  Returns the node (or None when the topology files are not there)."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU front end on the host cores.  The reference itself cannot be
    built here (needs OpenCV C++, Eigen, glog, g2o — see DESIGN.md), so this is the oracle port."""
    if rank != 0:
        return
    cores = host_cores()
    frames = make_pairs(max(cores, 8), seed=1)
    ppt = 1
    for _ in range(args.warmup):
        cpu_frontend_threads(frames, cores, ppt)
    tot_f = tot_k = 0; tot_t = 0.0
    for _ in range(args.steps):
        f, k, dt = cpu_frontend_threads(frames, cores, ppt)
        tot_f += f; tot_k += k; tot_t += dt
    fps = tot_f / tot_t
    line = {
        "impl": "reference", "metric": "frontend_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "keypoints_per_sec": tot_k / tot_t,
        "config": {"workload": f"KITTI-shaped stereo {W}x{H} pair, {NLEVELS} levels x1.2, {NFEAT} keypoints/image, "
                               "left-right windowed Hamming matching (BASELINE.json configs[1])",
                   "frames_per_step": 2 * ppt * cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": f"{2 * ppt * cores} frames ({ppt} stereo pair per thread) per step, {args.steps} steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=512, help="stereo pairs per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-contexts", type=int, default=2, help="contexts (host threads) that alternate in the e2e loop")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ORB path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    from nav24_b200 import capi
    P = args.pairs
    F = 2 * P
    ctx = capi.OrbContext(NFEAT, device=local)
    cap = None
    grid = capi.grid_for(W, H)
    pairs = np.stack([np.arange(0, F, 2), np.arange(1, F, 2)], 1).astype(np.int32)

    # two distinct input sets so that consecutive steps never see their inputs in L2
    sets_host = [make_pairs(P, seed=1000 * rank + 17 + s) for s in range(2)]
    dsets = []
    for fr in sets_host:
        padded = np.zeros((F, H, PITCH), np.uint8); padded[:, :, :W] = fr
        dptr = capi.C.c_void_p()
        assert capi.lib().nav24_device_alloc(padded.nbytes, capi.C.byref(dptr)) == 0
        assert capi.lib().nav24_memcpy_h2d(dptr, padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0
        dsets.append(dptr.value)

    def step_resident(i):      # fused detect + left-right matching, asynchronous
        ctx.detect_match_device(dsets[i % 2], F, W, H, PITCH, PITCH * H, pairs, grid)

    def step_serial(i):        # same work on ONE stream, un-overlapped: gives clean per-stage CUDA-event times
        ctx.detect_device(dsets[i % 2], F, W, H, PITCH, PITCH * H)
        ctx.match_window_frames_async(pairs, grid)

    sampler = ClockSampler(local)
    sampler.start()             # runs through warm-up and the timed region (nvidia-smi needs ~0.1 s to deliver its first line)
    for i in range(args.warmup):
        step_resident(i)
    ctx.sync()
    cap = ctx.max_keypoints()
    launches0 = ctx.launch_count()
    barrier()
    ctx.timer_start()
    for i in range(args.steps):
        step_resident(i)
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    n, mono, _, _ = ctx.fetch(F, want_data=False)
    m, nm = ctx.match_fetch(P, want_matches=False)
    # stage times (and the roofline's kernel time): the same steps again on one stream, nothing overlapped
    step_serial(0); ctx.sync()
    ctx.stage_ms_sum(reset=True)
    ctx.timer_start()
    for i in range(args.steps):
        step_serial(i)
    ms_serial = ctx.timer_stop()
    stage, calls = ctx.stage_ms_sum(reset=True)
    kp_per_frame = float(n.mean())
    raw_per_frame = float(np.mean([sum(ctx.L.nav24_orb_get_raw_keys(ctx.h, f, l, None, 0) for l in range(NLEVELS))
                                   for f in range(0, F, max(1, F // 8))]))

    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    fps = world * F * args.steps / (ms_max * 1e-3)

    # ---- end to end through the C ABI with pinned host buffers --------------------------------------------
    h_frames = [capi.pinned_empty((F, H, W)) for _ in range(2)]
    for hb, fr in zip(h_frames, sets_host):
        hb[...] = fr
    # Two contexts on the same GPU, one host thread each, take the steps alternately: every call is synchronous (it
    # returns with its results in host memory), so one call's fill (first copy in) and drain (last kernels, last copy
    # out) overlap the other call's steady state — how a streaming caller uses a synchronous batched API.  The C calls
    # release the GIL.  (--e2e-contexts 1: one context, strictly one call after the other.)
    n_ctx = max(1, args.e2e_contexts)
    ctxs = [ctx] + [capi.OrbContext(NFEAT, device=local) for _ in range(n_ctx - 1)]
    outs = [(capi.pinned_empty((F, cap), capi.KP_DTYPE), capi.pinned_empty((F, cap, 32)), capi.pinned_empty((P, cap), np.int32))
            for _ in range(n_ctx)]

    def step_e2e(c, i):
        k, d, m = outs[c]
        nn, mm, _, _, mt, nmt = ctxs[c].detect_match_batch(h_frames[i % 2], pairs, grid, cap=cap, kps=k, desc=d, matches=m)
        return nn, nmt

    def run_e2e(n_steps):
        errs = []

        def work(c):
            try:
                for i in range(c, n_steps, n_ctx):
                    step_e2e(c, i)
            except BaseException as e:      # a failed call must fail the bench, not shorten the timed region
                errs.append(e)
        if n_ctx == 1:
            work(0)
        else:
            ths = [threading.Thread(target=work, args=(c,)) for c in range(n_ctx)]
            for th in ths:
                th.start()
            for th in ths:
                th.join()
        if errs:
            raise errs[0]

    run_e2e(max(args.warmup, n_ctx))
    barrier()
    t0 = time.perf_counter()
    run_e2e(args.steps)
    e2e_s = time.perf_counter() - t0
    barrier()
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_fps = world * F * args.steps / float(t.item())
    h2d = F * H * W
    d2h = F * cap * (28 + 32) + P * cap * 4 + F * 8 + P * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of pyramid+FAST (the HBM-bound group BASELINE.json's metric names) -----------------------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    pix = level_pixels(H, W, NLEVELS)
    alg_bytes_frame = pix + 12.0 * raw_per_frame            # B_pyrFAST
    pf_ms = float(stage[0] + stage[1]) / max(calls, 1)      # per step
    achieved = alg_bytes_frame * F / (pf_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "pyramid (resize_kernel x7) + FAST (fast_band_kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of the 7 resize launches + the FAST launch from the ncu
                # --set full capture profiles/r01_v9_ncu_full_summary.md (870.8 MB per 256 frames), scaled to this batch
                "traffic": NCU_DRAM_BYTES_PER_FRAME * F, "traffic_source": "profiles/r01_v9_ncu_full_summary.md",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_frame * F,
                "avg_ms_per_step": pf_ms,
                "timing": "CUDA events on the library's stream around the pyramid and FAST launches, measured in a second pass "
                          "of the same steps issued on ONE stream (in the timed region two streams overlap chunks, which "
                          "would smear per-kernel times)", "serial_ms_per_step": ms_serial / args.steps}
    line = {
        "metric": "frontend_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "keypoints_per_sec": fps * kp_per_frame, "keypoints_per_frame": kp_per_frame,
        "matches_per_pair": float(nm.mean()),
        "config": {"workload": f"KITTI-shaped stereo {W}x{H} pair, {NLEVELS} levels x1.2, {NFEAT} keypoints/image, "
                               "left-right windowed Hamming matching (BASELINE.json configs[1])",
                   "pairs_per_step_per_gpu": P, "frames_per_step_per_gpu": F, "parallelism": f"sequences sharded x{world}, no collective",
                   "pipeline": "fused detect+match; resident: one launch set per step; host buffers: chunks of <= 64 frames (short first and last chunks), H2D / three compute streams / D2H overlapped, one host synchronisation per call", "e2e_contexts": max(1, args.e2e_contexts), "host_numa_node": numa,
                   "l2": f"two input sets alternate; per-step working set {(F * (pix * 2 + H * PITCH)) / 1e6:.0f} MB > 126 MB L2"},
        "stage_ms_per_step": {"pyramid": float(stage[0]) / max(calls, 1), "fast": float(stage[1]) / max(calls, 1),
                              "quadtree_order": float(stage[2]) / max(calls, 1),
                              "blur_orient_desc": float(stage[3]) / max(calls, 1),
                              "detect_total": float(stage[4]) / max(calls, 1)},
        "roofline": roofline,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline uses every host core, not only the GPU's NUMA node
        cores = host_cores()
        f1, k1, dt1 = cpu_frontend_threads(sets_host[0], 1, 4)
        per_thread = max(1, int(round(12.0 / (dt1 / 4))))
        per_thread = min(per_thread, 64)
        f, k, dt = cpu_frontend_threads(sets_host[0], cores, per_thread)
        line["cpu_baseline"] = {"value": f / dt, "unit": "frames/s", "cores": cores, "kind": "port",
                                "value_1core": f1 / dt1,
                                "sample": f"{f} frames of the same workload ({per_thread} stereo pairs per thread on {cores} "
                                          f"threads, {dt:.1f} s); 1-core figure from {f1} frames"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
