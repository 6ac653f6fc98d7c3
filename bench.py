#!/usr/bin/env python
"""bench.py — front-end throughput of the B200 ORB path on BASELINE.json's KITTI-shaped stereo workload.

One step = one pass of the hot path over one batch of synthetic stereo pairs: detect (pyramid, FAST
cells, quadtree, orientation, blur, rBRIEF) on every image of the batch + left-right windowed Hamming
matching of every pair.  `value` is measured with the frames already resident in HBM (CUDA events on
the library's own stream), `e2e` through the C-ABI calls with pinned HOST buffers (H2D of the frames and
D2H of keypoints, descriptors and matches inside the timed region).  Sequences shard across GPUs with no
collective (weak scaling: every rank processes its own batch).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs P] [--impl b200|reference]
                  [--workload stereo|seq64] [--batch-sweep]

--impl reference   the reference's own CPU front end (oracle/_ref: FtDtOrbSlam / FtAssocOrbSlam compiled from
                   /root/reference, else the oracle port) on all host cores, same workload.
--workload seq64   BASELINE.json configs[4]: 64 independent KITTI-shaped sequences x 100 frames, sharded seq % N
                   (nav24_b200/shard.py), frame t matched against the device-resident frame t-1 (strong scaling).
--batch-sweep      adds `batch_sweep`: pyramid+FAST fraction of the HBM peak vs frames per launch (1 ... 1024, and 4K).
"""
import argparse
import glob
import json
import os
import re
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NFEAT, NLEVELS = 376, 1241, 2000, 8
PITCH = 1280                      # 16-byte aligned row pitch of the device-resident frames
WORKLOAD = (f"KITTI-shaped stereo {W}x{H} pair, {NLEVELS} levels x1.2, {NFEAT} keypoints/image, "
            "left-right windowed Hamming matching (BASELINE.json configs[1])")


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


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------------------------
# CPU reference arm (the one place besides tests/ and smoke() that executes oracle/)
# ---------------------------------------------------------------------------------------------------------------
class CpuFrontEnd:
    """The reference's CPU front end (detect L, detect R, matchV) on `n_threads` PERSISTENT host threads: detector
    objects and threads are created once, outside every timed step.  kind = "reference": oracle/_ref, the reference's own
    FtDtOrbSlam / FtAssocOrbSlam / FeatureGrid compiled from /root/reference (pixel primitives: the oracle's scalar
    restatements of the OpenCV routines); kind = "port": the oracle restatement alone.  The ctypes calls release the GIL."""

    def __init__(self, n_threads, nfeat=NFEAT):
        from oracle import orb_oracle as oo
        self.oo = oo
        self.n = n_threads
        self.rl = None
        try:
            from oracle import ref_lib as rl
            if rl.available():
                self.rl = rl
        except Exception:
            self.rl = None
        self.kind = "reference" if self.rl else "port"
        oo.lib()
        if self.rl:
            self.det = [self.rl.RefOrb(nfeat) for _ in range(n_threads)]
            k = np.zeros(1, oo.KP_DTYPE); u = np.zeros((1, 2), np.float32); d = np.zeros((1, 32), np.uint8)
            self.rl.match_window(k, u, d, k, u, d, W, H)          # FeatureGrid::setImageBounds, once per process
        else:
            self.det = [oo.OrbOracle(nfeat) for _ in range(n_threads)]
            self.grid = oo.grid_for(W, H)
        self.pool = ThreadPoolExecutor(max_workers=n_threads)

    def _work(self, t, frames, first_pair, n_pairs):
        det, n_avail, nk = self.det[t], len(frames) // 2, 0
        for i in range(n_pairs):
            p = (first_pair + i) % n_avail
            _, k1, d1 = det.detect(frames[2 * p]); _, k2, d2 = det.detect(frames[2 * p + 1])
            ud1 = np.stack([k1["x"], k1["y"]], 1); ud2 = np.stack([k2["x"], k2["y"]], 1)
            if self.rl:
                self.rl.match_window(k1, ud1, d1, k2, ud2, d2, 0, 0)
            else:
                self.oo.match_window(k1, ud1, d1, k2, ud2, d2, self.grid)
            nk += len(k1) + len(k2)
        return nk

    def step(self, frames, pairs_per_thread):
        """One step: every thread processes pairs_per_thread stereo pairs.  Returns (frames, keypoints, seconds)."""
        t0 = time.perf_counter()
        futs = [self.pool.submit(self._work, t, frames, t * pairs_per_thread, pairs_per_thread) for t in range(self.n)]
        nk = sum(f.result() for f in futs)
        return 2 * pairs_per_thread * self.n, nk, time.perf_counter() - t0

    def close(self):
        self.pool.shutdown()


def primitive_times(frame, reps=3):
    """Milliseconds per KITTI frame, one thread, of the pixel primitives alone (pyramid resize chain, FAST(20) over whole
    levels, 7x7 blur of every level): the oracle's scalar restatements vs the OpenCV-SIMD library the real reference
    links (cv2, when it imports).  The reference's remaining cost (cell loop, quadtree, orientation, descriptors,
    matching) is its own scalar C++ in both cases."""
    from oracle import orb_oracle as oo
    sizes = []
    s = np.float32(1.0)
    for _ in range(NLEVELS):
        inv = np.float32(1.0) / s
        sizes.append((int(np.rint(np.float32(W) * inv)), int(np.rint(np.float32(H) * inv))))
        s = np.float32(float(s) * float(np.float32(1.2)))

    def run(resize, fast, blur):
        best = [1e9, 1e9, 1e9]
        for _ in range(reps):
            t0 = time.perf_counter()
            lv = [frame]
            for (w, h) in sizes[1:]:
                lv.append(resize(lv[-1], w, h))
            t1 = time.perf_counter()
            for a in lv:
                fast(a)
            t2 = time.perf_counter()
            for a in lv:
                blur(a)
            t3 = time.perf_counter()
            best = [min(b, v) for b, v in zip(best, (t1 - t0, t2 - t1, t3 - t2))]
        return {"pyramid": 1e3 * best[0], "fast": 1e3 * best[1], "blur": 1e3 * best[2], "total": 1e3 * sum(best)}

    out = {"port_scalar": run(oo.resize_u8, lambda a: oo.fast_u8(a, 20), oo.gauss7_u8)}
    try:
        import cv2
        cv2.setNumThreads(1)
        det = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
        out["cv2_simd_1thread"] = run(lambda a, w, h: cv2.resize(a, (w, h), interpolation=cv2.INTER_LINEAR), det.detect,
                                      lambda a: cv2.GaussianBlur(a, (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101))
        out["cv2_version"] = cv2.__version__
    except Exception as e:                       # cv2 is test infrastructure of the build container; it may be absent on the box
        out["cv2_simd_1thread"] = None
        out["cv2_note"] = f"cv2 not importable here: {type(e).__name__}"
    return out


def cpu_baseline(frames, seconds=12.0):
    """cpu_baseline object of the bench line: the CPU front end on all host cores over a bounded sample of the workload."""
    cores = host_cores()
    one = CpuFrontEnd(1)
    one.step(frames, 1)
    f1, k1, dt1 = one.step(frames, 4)
    one.close()
    fe = CpuFrontEnd(cores)
    fe.step(frames, 1)                                            # warm-up (page-in, thread start)
    per_thread = int(min(64, max(4, round(seconds / (dt1 / 4)))))
    f, k, dt = fe.step(frames, per_thread)
    fe.close()
    prim = primitive_times(frames[0])
    out = {"value": f / dt, "unit": "frames/s", "cores": cores, "kind": fe.kind, "value_1core": f1 / dt1,
           "keypoints_per_sec": k / dt,
           "sample": f"{f} frames of the same workload ({per_thread} stereo pairs per thread on {cores} persistent "
                     f"threads, {dt:.1f} s); 1-core figure from {f1} frames",
           "implementation": ("oracle/_ref: the reference's own FtDtOrbSlam / FtAssocOrbSlam / FeatureGrid (compiled from "
                              "/root/reference), pixel primitives = scalar restatements of the OpenCV routines"
                              if fe.kind == "reference" else "oracle/orb_oracle.cpp (scalar restatement)"),
           "primitives_ms_per_frame_1thread": prim}
    if prim.get("cv2_simd_1thread"):
        # what the reference would do with OpenCV's SIMD primitives instead of the scalar ones: same non-primitive remainder
        ms_frame = 1e3 * dt1 / f1
        est = ms_frame - prim["port_scalar"]["total"] + prim["cv2_simd_1thread"]["total"]
        out["opencv_simd_estimate"] = {
            "frames_per_sec_1core": 1e3 / est, "frames_per_sec_all_cores": cores * 1e3 / est * (f / dt) / (cores * f1 / dt1),
            "how": "1-core ms/frame minus the scalar primitives' time plus cv2's time for the same primitives; all-core figure "
                   "scaled by the measured multi-thread efficiency"}
    return out


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (rank 0 only)."""
    if rank != 0:
        return
    cores = host_cores()
    ppt = 4
    frames = make_pairs(max(ppt * cores, 8), seed=1)
    fe = CpuFrontEnd(cores)
    for _ in range(max(1, args.warmup)):
        fe.step(frames, ppt)
    tot_f = tot_k = 0; tot_t = 0.0
    for _ in range(args.steps):
        f, k, dt = fe.step(frames, ppt)
        tot_f += f; tot_k += k; tot_t += dt
    fe.close()
    fps = tot_f / tot_t
    line = {
        "impl": "reference", "metric": "frontend_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "keypoints_per_sec": tot_k / tot_t,
        "config": {"workload": WORKLOAD, "frames_per_step": 2 * ppt * cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": fe.kind,
                         "sample": f"{2 * ppt * cores} frames ({ppt} stereo pairs per persistent thread, {cores} threads) per step, "
                                   f"{args.steps} steps; detectors and threads built outside the timed steps"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------------------------
def numa_info(torch, local):
    """NUMA node of this rank's GPU (None when the box exposes no topology: a single-node VM) and the node count."""
    nodes = len(glob.glob("/sys/devices/system/node/node[0-9]*"))
    try:
        pr = torch.cuda.get_device_properties(local)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None, nodes
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus and nodes > 1:          # pin BEFORE any pinned buffer is allocated (first touch places the pages)
            os.sched_setaffinity(0, cpus)
        return node, nodes
    except Exception:
        return None, nodes


def newest_ncu_summary():
    """(path, frames per launch, {kernel: [rows]}) of the newest profiles/r*_ncu_full_summary.md, or None.  The summary
    starts with `<!-- frames_per_launch: N -->` (tools/ncu_summary.py); round-1 files without it were 256 frames."""
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_full_summary.md")),
                   key=lambda p: [int(x) for x in re.findall(r"\d+", os.path.basename(p))])
    if not files:
        return None
    path = files[-1]
    txt = open(path).read()
    m = re.search(r"frames_per_launch:\s*(\d+)", txt)
    frames = int(m.group(1)) if m else 256
    rows, hdr = {}, None
    for line in txt.splitlines():
        if not line.startswith("|") or line.startswith("|---"):
            continue
        cells = [c.strip() for c in line.strip().strip("|").split("|")]
        if cells[0] == "kernel":
            hdr = cells; continue
        if hdr:
            rows.setdefault(cells[0], []).append(dict(zip(hdr, cells)))
    return path, frames, rows


def num(s):
    m = re.match(r"[-+]?[\d.]+", s)
    return float(m.group(0)) if m else 0.0


def pyrfast_from_summary(pix_per_frame):
    """DRAM traffic per frame (bytes) and executed thread instructions per pixel of pyramid + FAST from the newest ncu
    --set full summary; (None, None, None) when no summary is committed."""
    s = newest_ncu_summary()
    if not s:
        return None, None, None
    path, frames, rows = s
    dram = inst = 0.0
    found = False
    for name, rr in rows.items():
        if "resize" in name or "fast_band" in name or "pyrfast" in name:      # (names are cut short in the summary: resize8_kerne...)
            found = True
            for r in rr:
                dram += (num(r.get("dram_rd", "0")) + num(r.get("dram_wr", "0"))) * 1e6
                inst += num(r.get("tinst", "0")) * 1e6 if "tinst" in r else num(r.get("inst", "0")) * 32e6
    if not found:
        return None, None, None
    return dram / frames, inst / frames / pix_per_frame, os.path.relpath(path, ROOT)


def pyrfast_warp_inst_per_frame():
    """Warp-level instructions pyramid + FAST issue per frame (newest ncu summary), or None: the numerator of the issue-slot
    roofline, the one that actually bounds the group."""
    s = newest_ncu_summary()
    if not s:
        return None
    _, frames, rows = s
    inst = sum(num(r.get("inst", "0")) * 1e6 for name, rr in rows.items() if "resize" in name or "fast_band" in name for r in rr)
    return inst / frames if inst > 0 else None


def issue_roofline(pf_ms, frames, clocks, sms=148):
    wi = pyrfast_warp_inst_per_frame()
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    if wi is None or pf_ms <= 0:
        return None
    achieved = wi * frames / (pf_ms * 1e-3) / 1e9
    peak = sms * 4 * mhz * 1e6 / 1e9
    return {"achieved": achieved, "peak": peak, "unit": "G warp-instructions/s", "frac": achieved / peak,
            "how": "warp instructions of the resize + FAST launches per frame (newest ncu --set full summary) x frames per step / "
                   "the group's CUDA-event time; peak = 148 SMs x 4 schedulers x SM clock under load"}


def copy_ceiling(torch, barrier, world, h2d_bytes, d2h_bytes, chunks=16, seconds=1.0):
    """Pure-copy ceiling of the host<->device path of THIS box with all ranks copying at once: pinned H2D of one step's
    frames in `chunks` pieces on one stream, pinned D2H of one step's results on another, no kernels.  Returns the
    aggregate GB/s over all ranks in each direction (bytes of all ranks / max-over-ranks wall time)."""
    import torch.distributed as dist
    hin = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True); din = torch.empty(h2d_bytes, dtype=torch.uint8, device="cuda")
    hout = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True); dout = torch.empty(d2h_bytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ci, co = (h2d_bytes + chunks - 1) // chunks, (d2h_bytes + chunks - 1) // chunks

    def one():
        for c in range(chunks):
            with torch.cuda.stream(s1):
                din[c * ci:(c + 1) * ci].copy_(hin[c * ci:(c + 1) * ci], non_blocking=True)
            with torch.cuda.stream(s2):
                hout[c * co:(c + 1) * co].copy_(dout[c * co:(c + 1) * co], non_blocking=True)
    one(); torch.cuda.synchronize()
    t0 = time.perf_counter(); one(); torch.cuda.synchronize()
    iters = int(max(2, min(200, seconds / max(time.perf_counter() - t0, 1e-4))))
    barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        one()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    barrier()
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    return world * iters * h2d_bytes / dt / 1e9, world * iters * d2h_bytes / dt / 1e9


def check_frames_against_oracle(o, oo, frames, n, mono, kps, desc):
    """Frames of the CUDA path against the oracle: counts mismatching frames (keypoint records byte-equal, monoIndex,
    descriptors within the 0.1 % tolerance)."""
    bad_frames = bad_desc = 0
    for f in range(len(frames)):
        mo, ko, do = o.detect(frames[f])
        ok = mono[f] == mo and n[f] == len(ko) and kps[f, :n[f]].tobytes() == ko.tobytes()
        if ok:
            nd = int((desc[f, :n[f]] != do).any(axis=1).sum())
            bad_desc += nd
            ok = nd <= 1e-3 * max(1, len(ko))
        bad_frames += 0 if ok else 1
    return bad_frames, bad_desc


def oracle_matches(oo, kps, desc, n, a, b):
    k1, k2 = kps[a, :n[a]], kps[b, :n[b]]
    return oo.match_window(k1, np.stack([k1["x"], k1["y"]], 1), desc[a, :n[a]], k2, np.stack([k2["x"], k2["y"]], 1),
                           desc[b, :n[b]], oo.grid_for(W, H))


def gather_parity(dist, world, part):
    parts = [part]
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, part)
    tot = {"frames": 0, "pairs": 0, "mismatches": 0, "descriptor_mismatches": 0}
    for p in parts:
        for k in tot:
            tot[k] += p[k]
    tot["ranks"] = world
    tot["against"] = "oracle/orb_oracle.cpp on the same frames, after the timed region (keypoints byte-equal, matches index-equal)"
    return tot


def run_seq64(args, torch, dist, rank, world, local, barrier):
    """BASELINE.json configs[4]: 64 independent KITTI-shaped sequences x 100 frames sharded across the GPUs of the box
    (nav24_b200.shard: seq % world), every frame matched against the device-resident previous frame of its sequence.
    Strong scaling: the total work is fixed.  After the timed region every rank checks two of its sequences against
    the oracle and the per-sequence digests are gathered on rank 0."""
    from nav24_b200 import capi, shard
    from nav24_b200.synth import sequence
    from oracle import orb_oracle as oo
    NSEQ, T = args.sequences, args.seq_frames
    mine = shard.sequences_for_rank(NSEQ, rank, world)
    S = len(mine)
    F = S * T
    ctx = capi.OrbContext(NFEAT, device=local)
    grid = capi.grid_for(W, H)
    pairs = np.array([(s * T + t - 1, s * T + t) for s in range(S) for t in range(1, T)], np.int32).reshape(-1, 2)
    P = len(pairs)
    dptr = capi.C.c_void_p()
    assert capi.lib().nav24_device_alloc(max(F, 1) * H * PITCH, capi.C.byref(dptr)) == 0
    host_seq = {}
    for i, s in enumerate(mine):      # seed = 1000 * seq + t: one canvas per sequence, frame t shifted by (t mod 60, 0)
        fr = sequence(H, W, 1000 * s + 7, T, step=(1, 0))
        if i < 2:
            host_seq[i] = fr
        padded = np.zeros((T, H, PITCH), np.uint8); padded[:, :, :W] = fr
        assert capi.lib().nav24_memcpy_h2d(capi.C.c_void_p(dptr.value + i * T * H * PITCH),
                                           padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0

    def step():
        if F:
            ctx.detect_match_device(dptr.value, F, W, H, PITCH, PITCH * H, pairs, grid)
    sampler = ClockSampler(local); sampler.start()
    for _ in range(args.warmup):
        step()
    ctx.sync()
    l0 = ctx.launch_count()
    barrier()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    # per-sequence summaries (frames, keypoints, matches, digest) and the oracle check of two sequences per rank
    summary, part = {}, {"frames": 0, "pairs": 0, "mismatches": 0, "descriptor_mismatches": 0}
    o = oo.OrbOracle(NFEAT)
    for i, s in enumerate(mine):
        n, mono, kps, desc = ctx.fetch_range(i * T, T)
        m, nm = ctx.match_fetch_range(i * (T - 1), T - 1)
        summary[s] = (T, int(n.sum()), int(nm.sum()), shard.digest(n, mono, *[kps[f, :n[f]] for f in range(T)],
                                                                  *[desc[f, :n[f]] for f in range(T)],
                                                                  *[m[q, :n[q]] for q in range(T - 1)]))
        if i in host_seq:
            bf, bd = check_frames_against_oracle(o, oo, host_seq[i], n, mono, kps, desc)
            bm = sum(0 if np.array_equal(m[q, :n[q]], oracle_matches(oo, kps, desc, n, q, q + 1)) else 1 for q in range(T - 1))
            part["frames"] += T; part["pairs"] += T - 1; part["mismatches"] += bf + bm; part["descriptor_mismatches"] += bd
    merged = shard.gather_summaries(summary)
    parity = gather_parity(dist, world, part)
    if rank != 0:
        return
    tot_frames = sum(v[0] for v in merged.values())
    fps = tot_frames * args.steps / (ms_max * 1e-3)
    kp = sum(v[1] for v in merged.values())
    line = {
        "metric": "frontend_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "keypoints_per_sec": fps * kp / max(tot_frames, 1),
        "config": {"workload": f"{NSEQ} independent KITTI-shaped sequences x {T} frames ({W}x{H}, {NFEAT} keypoints), frame t "
                               "matched against the device-resident frame t-1 (BASELINE.json configs[4])",
                   "parallelism": f"sequences sharded seq % {world} (nav24_b200/shard.py), no collective",
                   "sequences_per_gpu": [len(shard.sequences_for_rank(NSEQ, r, world)) for r in range(world)],
                   "l2": f"per-step working set {(F * (level_pixels(H, W) * 2 + H * PITCH)) / 1e6:.0f} MB per GPU > 126 MB L2"},
        "sequences": len(merged), "frames_per_step": tot_frames, "matches_per_step": sum(v[2] for v in merged.values()),
        "digest_of_digests": shard.digest(np.array([merged[s][3] for s in sorted(merged)], np.uint64)),
        "parity_checked": parity, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)


def batch_sweep(capi, local, peak):
    """pyramid+FAST fraction of the HBM peak vs frames per launch: device-resident detect only."""
    out = []
    for (h, w, nf, batches) in ((H, W, NFEAT, (1, 8, 64, 256, 1024)), (2160, 3840, 8000, (1, 8, 32))):
        from nav24_b200.synth import synth
        base = synth(h, w, 5)
        pitch = (w + 127) // 128 * 128
        ctx = capi.OrbContext(nf, device=local)
        try:
            for B in batches:
                padded = np.zeros((B, h, pitch), np.uint8); padded[:, :, :w] = base
                dptr = capi.C.c_void_p()
                assert capi.lib().nav24_device_alloc(padded.nbytes, capi.C.byref(dptr)) == 0
                assert capi.lib().nav24_memcpy_h2d(dptr, padded.ctypes.data_as(capi.C.c_void_p), padded.nbytes) == 0
                for _ in range(3):
                    ctx.detect_device(dptr.value, B, w, h, pitch, pitch * h)
                ctx.sync(); ctx.stage_ms_sum(reset=True)
                reps = max(3, min(50, 2048 // B))
                for _ in range(reps):
                    ctx.detect_device(dptr.value, B, w, h, pitch, pitch * h)
                ctx.sync()
                st, calls = ctx.stage_ms_sum(reset=True)
                raw = float(np.mean([sum(ctx.L.nav24_orb_get_raw_keys(ctx.h, f, l, None, 0) for l in range(NLEVELS))
                                     for f in range(0, B, max(1, B // 4))]))
                pf_ms = float(st[0] + st[1]) / calls
                alg = (level_pixels(h, w) + 12.0 * raw) * B
                out.append({"shape": f"{w}x{h}", "frames_per_launch": B, "pyramid_fast_ms": pf_ms,
                            "detect_ms": float(st[4]) / calls, "frames_per_sec": B / (float(st[4]) / calls * 1e-3),
                            "achieved_gbs": alg / (pf_ms * 1e-3) / 1e9, "frac": alg / (pf_ms * 1e-3) / 1e9 / peak})
                capi.lib().nav24_device_free(dptr)
        finally:
            ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--pairs", type=int, default=512, help="stereo pairs per step per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="stereo", choices=["stereo", "seq64"])
    ap.add_argument("--sequences", type=int, default=64)
    ap.add_argument("--seq-frames", type=int, default=100)
    ap.add_argument("--batch-sweep", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-copy-ceiling", action="store_true")
    ap.add_argument("--e2e-contexts", type=int, default=2, help="contexts (host threads) that alternate in the e2e loop")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the ORB path has no CPU fallback")
    torch.cuda.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa, numa_nodes = numa_info(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    if args.workload == "seq64":
        run_seq64(args, torch, dist, rank, world, local, barrier)
        if world > 1:
            dist.destroy_process_group()
        return

    from nav24_b200 import capi
    P = args.pairs
    F = 2 * P
    ctx = capi.OrbContext(NFEAT, device=local)
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

    # ---- hardware-side correctness record of THIS rank's resident results (every N): first and last stereo pair of the
    # last timed step against the oracle, gathered over the ranks
    from oracle import orb_oracle as oo          # checker only, after the timed region
    last = sets_host[(args.steps - 1) % 2]
    part = {"frames": 0, "pairs": 0, "mismatches": 0, "descriptor_mismatches": 0}
    o = oo.OrbOracle(NFEAT)
    for p in sorted({0, P - 1}):
        nn, mm, kk, dd = ctx.fetch_range(2 * p, 2)
        mt, _ = ctx.match_fetch_range(p, 1)
        bf, bd = check_frames_against_oracle(o, oo, last[2 * p:2 * p + 2], nn, mm, kk, dd)
        bm = 0 if np.array_equal(mt[0, :nn[0]], oracle_matches(oo, kk, dd, nn, 0, 1)) else 1
        part["frames"] += 2; part["pairs"] += 1; part["mismatches"] += bf + bm; part["descriptor_mismatches"] += bd
    parity = gather_parity(dist, world, part)

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
        k, d, mm = outs[c]
        ctxs[c].detect_match_batch(h_frames[i % 2], pairs, grid, cap=cap, kps=k, desc=d, matches=mm)

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
    e2e = {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "h2d_gbs_all_ranks": e2e_fps * H * W / 1e9, "d2h_gbs_all_ranks": e2e_fps * d2h / F / 1e9}
    if not args.no_copy_ceiling:
        for c in ctxs[1:]:
            c.close()
        hg, dg = copy_ceiling(torch, barrier, world, h2d, d2h)
        e2e["host_copy_ceiling_gbs"] = {"h2d": hg, "d2h": dg, "how": "all ranks at once, pinned buffers, the step's bytes in 16 pieces "
                                        "per direction on two streams, no kernels (torch copies; max-over-ranks wall time)"}
        e2e["host_copy_ceiling_frames_per_sec"] = hg * 1e9 / (H * W)
        e2e["fraction_of_copy_ceiling"] = e2e_fps / (hg * 1e9 / (H * W))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of pyramid+FAST (the group BASELINE.json's metric names) ----------------------------------
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    pix = level_pixels(H, W, NLEVELS)
    alg_bytes_frame = pix + 12.0 * raw_per_frame            # B_pyrFAST
    pf_ms = float(stage[0] + stage[1]) / max(calls, 1)      # per step
    achieved = alg_bytes_frame * F / (pf_ms * 1e-3) / 1e9
    traffic_frame, inst_px, src = pyrfast_from_summary(pix)
    roofline = {"bound": "hbm", "kernel": "pyramid (resize8_kernel / resize_kernel, 7 levels) + FAST (fast_band_kernel)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # measured limiter (ncu): the group is bound by instruction issue on the integer pipes, not by HBM; the HBM
                # roofline is the one BASELINE.json's metric asks the group to be reported against
                "limiter": "instruction issue (integer ALU / LSU pipes), DRAM < 25 % busy — see profiles/",
                "inst_per_px": inst_px,
                # the roofline that does bound the group: warp instructions issued (ncu summary) / (SMs x 4 schedulers x clock)
                "issue": issue_roofline(pf_ms, F, clocks),
                "traffic": traffic_frame * F if traffic_frame is not None else None, "traffic_source": src,
                "traffic_over_algorithmic": (traffic_frame / alg_bytes_frame) if traffic_frame is not None else None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes_frame * F,
                "avg_ms_per_step": pf_ms,
                "timing": "CUDA events on the library's stream around the pyramid and FAST launches, measured in a second pass "
                          "of the same steps issued on ONE stream (in the timed region the streams overlap chunks, which "
                          "would smear per-kernel times)", "serial_ms_per_step": ms_serial / args.steps}
    line = {
        "metric": "frontend_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "keypoints_per_sec": fps * kp_per_frame, "keypoints_per_frame": kp_per_frame,
        "matches_per_pair": float(nm.mean()),
        "config": {"workload": WORKLOAD,
                   "pairs_per_step_per_gpu": P, "frames_per_step_per_gpu": F, "parallelism": f"sequences sharded x{world}, no collective",
                   "pipeline": "fused detect+match; resident: one launch set per half of the step's frames, the halves on two streams (batches >= 512 frames), resize / FAST / quadtree launched as programmatic dependents; host buffers: chunks of <= 64 frames (short first and last chunks), H2D / three compute streams / D2H overlapped, one host synchronisation per call",
                   "e2e_contexts": n_ctx, "host_numa_node": numa, "host_numa_nodes": numa_nodes, "host_cores": len(all_cpus),
                   "l2": f"two input sets alternate; per-step working set {(F * (pix * 2 + H * PITCH)) / 1e6:.0f} MB > 126 MB L2"},
        "stage_ms_per_step": {"pyramid": float(stage[0]) / max(calls, 1), "fast": float(stage[1]) / max(calls, 1),
                              "quadtree_order": float(stage[2]) / max(calls, 1),
                              "blur_orient_desc": float(stage[3]) / max(calls, 1),
                              "detect_total": float(stage[4]) / max(calls, 1)},
        "roofline": roofline,
        "e2e": e2e,
        "parity_checked": parity,
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if args.batch_sweep and world == 1:
        line["batch_sweep"] = batch_sweep(capi, local, peak)
    if world == 1 and not args.no_cpu_baseline:
        os.sched_setaffinity(0, all_cpus)      # the CPU baseline uses every host core, not only the GPU's NUMA node
        line["cpu_baseline"] = cpu_baseline(sets_host[0])
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
