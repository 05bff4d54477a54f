"""GPU probe: thread-per-query vs warp-per-query kNN by list length (B2_KNN_TRACE=1 prints the search kernel's device time)."""
import sys
import numpy as np
sys.path.insert(0, ".")
import dataset_pipeline_b200 as b2
from dataset_pipeline_b200 import synth
rng = np.random.default_rng(0)
b2.estimate_normals(rng.uniform(0, 1, (100000, 3)).astype(np.float32), 20)
s, _, _ = synth.room_scan(0, 1600, 640, seed=20)
for k in [int(a) for a in sys.argv[1:]]:
    b2.estimate_normals(s, k)
