#!/bin/bash
# compute-sanitizer (memcheck, racecheck, synccheck) over one small end-to-end invocation of EVERY kernel of the library:
# detect + match (host pipeline and device-resident), undistort, brute-force matcher, colour ingest ring, two-view scoring,
# the dense FAST path (noise), the global-table quadtree path (x5 feature mode) and the 1024-thread small-batch variants.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
from nav24_b200 import capi
from nav24_b200.synth import sequence, synth
fr = sequence(260, 340, 5, 4, step=(3, 1), lowtex=True)
ctx = capi.OrbContext(300)
n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, [(0, 1), (2, 3), (0, 3)], capi.grid_for(340, 260))
print('kps', n, 'matches', nm)
ctx.set_camera(capi.Camera.make(capi.CAM_RADTAN, [300, 300, 170, 130], [-0.2, 0.05, 0.001, 0.0005]))
n, mono, kps, desc, m, nm = ctx.detect_match_batch(fr, [(0, 1)], capi.grid_for(340, 260))
print('ud', ctx.fetch_undistorted(4)[0, :2])
ctx.set_camera(None)
rng = np.random.default_rng(1)
noise = rng.integers(0, 256, (300, 400), dtype=np.uint8)
c2 = capi.OrbContext(1000, raw_keys_per_kpx=250)
print('noise', len(c2.detect(noise)[1]))
big = synth(376, 1241, 3)
c3 = capi.OrbContext(2000)
print('kitti', len(c3.detect(big)[1]))
c3.set_num_features(10000)
print('kitti x5', len(c3.detect(big)[1]))
many = np.stack([big] * 40)                          # 320 (level, frame) CTAs: the 256-thread quadtree with shared-memory node tables
c3.set_num_features(2000)
print('batch40', int(c3.detect_batch(many)[0].sum()))
print('bf', ctx.match_bf_knn2(desc[0, :n[0]], desc[1, :n[1]])[4].sum())
ring = capi.IngestRing(ctx, 341, 260, 3, 3)
for k in range(3):
    ring.slot(k)[...] = rng.integers(0, 256, (260, 341, 3), dtype=np.uint8) // 4 * 4
print('ring', ring.detect_match(0, 3, [(0, 1)], capi.grid_for(341, 260))[0])
print('ring1', ring.detect_match(1, 1)[0])
x1 = rng.uniform(0, 700, (777, 2)).astype(np.float32); x2 = x1 + np.float32(2.5)
H = np.tile(np.array([1, 0, 2.5, 0, 1, 2.5, 0, 0, 1], np.float32), (9, 1)); Hi = np.tile(np.array([1, 0, -2.5, 0, 1, -2.5, 0, 0, 1], np.float32), (9, 1))
F = np.tile(np.array([0, 0, 2.5, 0, 0, -2.5, -2.5, 2.5, 0], np.float32), (9, 1))
r = ctx.two_view_score(x1, x2, H, Hi, F)
print('twoview', r['best_h'], r['best_f'], float(r['score_h'][0]))
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do
    compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
    echo "== $tool: exit $?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|kps|noise|kitti|batch40|bf|ud|ring|twoview" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -24
done 2>&1 | tee gpurun_out/sanitize_summary.txt
