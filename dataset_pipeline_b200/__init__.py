"""dataset_pipeline_b200 — B200-native (sm_100a) hot paths of ETH3D/dataset-pipeline behind a C ABI.

Only the two data-parallel hot paths are here: Path A multi-scan point-to-plane ICP + kNN normals,
Path B photometric image<->scan alignment. Everything runs in libeth3d_b200.so (hand-written CUDA); this package is
the thin host-side mirror of the reference's class interfaces.
"""
from . import _lib  # noqa: F401
from .icp import Comm, PointToPlaneICP, find_correspondences  # noqa: F401
from .normals import NormalEstimationTwoPassOMP, estimate_normals, estimate_normals_radius  # noqa: F401
from .registration import Registration  # noqa: F401
from .multiscale import CreateMultiScalePointCloud, DeterminePointNeighbors, MergeClosePoints  # noqa: F401
from .scan_aligner import align_scans, scale_schedule  # noqa: F401
from .cleaner import LocalStatisticalOutlierRemoval, clean_point_cloud, create_splats, mesh_squared_distance  # noqa: F401
