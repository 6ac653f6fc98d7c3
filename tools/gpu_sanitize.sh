#!/bin/bash
# compute-sanitizer memcheck over one small end-to-end invocation (detect + match + undistort + bf matcher)
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
print('bf', ctx.match_bf_knn2(desc[0, :n[0]], desc[1, :n[1]])[4].sum())
PY
compute-sanitizer --tool ${SAN_TOOL:-memcheck} --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitize.log 2>&1
echo "exit $?" >> gpurun_out/sanitize.log
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard|exit|kps|noise|kitti|bf|ud" gpurun_out/sanitize.log | sort | uniq -c | sort -rn | head -30
