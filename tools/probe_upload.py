import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B
import dataset_pipeline_b200 as b2
B.ensure_scans(range(8), 5000, 2000)
start, gt = B.scene_poses(8)
clouds = []
for i in range(8):
    xyz, nrm = B.load_scan(i, 5000, 2000)
    px = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True); pn = torch.empty(nrm.shape, dtype=torch.float32, pin_memory=True)
    px.numpy()[:] = xyz; pn.numpy()[:] = nrm; clouds.append((px, pn))
for hint in (0.0, 0.01, 0.0, 0.01):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    g = b2.PointToPlaneICP(index_distance_hint=hint)
    ts = []
    for (px, pn), T in zip(clouds, gt):
        t = time.perf_counter(); g.AddPointCloud(px.numpy(), pn.numpy(), T); ts.append(1e3 * (time.perf_counter() - t))
    t1 = time.perf_counter(); g.Run(0.01, 0, 1, 1e-10, False); t2 = time.perf_counter()
    st = g.stats(); t3 = time.perf_counter(); g.close(); t4 = time.perf_counter()
    print("hint %.2f: create+add %.1f ms [%s]  run %.1f ms (index_build %.2f, device total %.1f)  destroy %.1f ms" % (
        hint, 1e3 * (t1 - t0), " ".join("%.1f" % v for v in ts), 1e3 * (t2 - t1), st["ms_index_build"], st["ms_total"], 1e3 * (t4 - t3)))
