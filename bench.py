#!/usr/bin/env python
"""Benchmark of the ICP hot path (BASELINE.json configs[1]): ICPScanAligner on 8 x 10M-point synthetic scans, point-to-plane.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A step = one ICP outer iteration = one `icp.Run(d, it, 1, thr)` as issued by icp_scan_aligner.cc:343 (global-frame transform,
all pair-direction correspondence searches, the complete inner LM loop of PointToPlaneICPImpl::compute, pose composition).

Which iteration. Both arms time the SAME, fully specified step: an outer iteration started from `--state` (default "gt": the
scanner poses of the synthetic scene, i.e. the refinement iterations an alignment spends most of its outer loop in; the poses
move, Run() does not report convergence). The B200 arm puts the poses back to that state before every timed step; the reference
arm executes that one iteration for real, once, on the full 8 x 10M scans (steps: 1) — its serial accumulate / cost loops take
minutes per outer iteration, and an iteration started from the 5 mm / 0.1 deg perturbed poses ("start", dozens of LM iterations)
would take it the better part of an hour. The B200 arm additionally reports, under "extra", the whole alignment from the
perturbed start (iterations 0..k until Run() reports convergence: per-iteration LM counts and times) and the post-convergence
iteration.
Prints ONE JSON line (rank 0). Data: synthetic room scans (dataset_pipeline_b200/synth), inputs far larger than L2.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ICP iterations/sec (8x10M-pt scans, point-to-plane, d=0.01)"
UNIT = "iterations/s"
CACHE = os.environ.get("B2_SYNTH_CACHE", "/tmp/b2_synth_cache")
THR = 1e-10


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scans", type=int, default=8)
    ap.add_argument("--scan-w", type=int, default=5000)
    ap.add_argument("--scan-h", type=int, default=2000)
    ap.add_argument("--d", type=float, default=0.01)
    ap.add_argument("--state", default="gt", choices=["gt", "start", "trajectory"],
                    help="pose state every timed step starts from: gt = the scene's scanner poses; start = the perturbed poses; "
                         "trajectory = free-running from the perturbed poses, back to them whenever Run() reports convergence")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the untimed alignment-from-start report")
    ap.add_argument("--no-secondary", action="store_true", help="skip the ImageRegistrator (Path B) numbers")
    return ap.parse_args()


def workload(args):
    return {"workload": "ICPScanAligner %dx%dM-pt synthetic room scans, point-to-plane, d=%g, 1 outer iteration per step, every step started from state '%s'"
                        % (args.scans, args.scan_w * args.scan_h // 1000000, args.d, args.state),
            "scans": args.scans, "points_per_scan": args.scan_w * args.scan_h, "state": args.state}


# --------------------------------------------------------------------------------------------------------------------
# inputs
# --------------------------------------------------------------------------------------------------------------------
def scan_path(i, W, H):
    return os.path.join(CACHE, "room_%dx%d_scan%d" % (W, H, i))


def ensure_scans(indices, W, H):
    """Generate (once per box) and cache the synthetic scans; returns nothing."""
    from dataset_pipeline_b200 import synth
    os.makedirs(CACHE, exist_ok=True)
    for i in indices:
        p = scan_path(i, W, H)
        if os.path.exists(p + ".xyz.npy") and os.path.exists(p + ".nrm.npy"):
            continue
        xyz, nrm, _ = synth.room_scan(i, W, H)
        np.save(p + ".xyz.tmp.npy", xyz); np.save(p + ".nrm.tmp.npy", nrm)
        os.replace(p + ".xyz.tmp.npy", p + ".xyz.npy"); os.replace(p + ".nrm.tmp.npy", p + ".nrm.npy")


def load_scan(i, W, H):
    p = scan_path(i, W, H)
    return np.load(p + ".xyz.npy"), np.load(p + ".nrm.npy")


def scene_poses(nscans):
    """(perturbed start poses, ground-truth poses) of the config-2 scene."""
    from dataset_pipeline_b200 import synth
    gts = []
    for i in range(nscans):
        T = np.eye(4); T[:3, :3] = synth.rot_xyz(0, 0, 0.35 * i); T[:3, 3] = synth.SCANNER_POSITIONS[i]
        gts.append(T)
    return synth.perturbed_poses(gts), [T.astype(np.float32) for T in gts]


def state_poses(args):
    start, gt = scene_poses(max(2, args.scans))
    return gt if args.state == "gt" else start


# --------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port (reference algorithm and parallel structure: OpenMP over pair-directions only, serial accumulate /
# cost loops) executing REAL outer iterations
# --------------------------------------------------------------------------------------------------------------------
def cpu_iteration(args, scan_ids, poses):
    """One real outer iteration of the oracle on the given scans (full size) from `poses`. Returns (seconds, stats, cores)."""
    from oracle import oracle as orc
    orc.build()
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    icp = orc.PointToPlaneICP(use_kdtree=True)
    for i in scan_ids:
        xyz, nrm = load_scan(i, args.scan_w, args.scan_h)
        icp.AddPointCloud(xyz, nrm, poses[i])
        del xyz, nrm
    t0 = time.perf_counter()
    icp.Run(args.d, 0, 1, THR, False)
    dt = time.perf_counter() - t0
    st = icp.stats()
    st["tries"] = [int(v) for v in icp.tries()]
    return dt, st, cores


def cpu_baseline_sample(args, counts):
    """cpu_baseline of the B200 line: a bounded sample = one REAL outer iteration of the oracle on scans 0 and 1 at full size from the
    same state (2 of the 56 pair-directions, the complete LM loop on their correspondences), scaled to the full step with the measured
    rates: search time x 56 / 2 over the cores OpenMP can use (it parallelises over pair-directions only), LM time x (passes x C) of the
    full step as counted by the B200 run of the identical step."""
    ensure_scans([0, 1], args.scan_w, args.scan_h)
    poses = state_poses(args)
    dt, st, cores = cpu_iteration(args, [0, 1], poses)
    ndirs = args.scans * (args.scans - 1)
    C_s = max(1, st["num_correspondences"])
    r_acc = st["t_acc"] / (C_s * max(1, st["inner_iterations"]))
    r_cost = st["t_cost"] / (C_s * max(1, st["lm_tries_total"]))
    t_search_dir = st["t_search"]                 # the two directions ran concurrently: wall time of one
    t_full = (st["t_transform"] / 2.0) * args.scans + t_search_dir * math.ceil(ndirs / min(cores, ndirs)) + \
        counts["C"] * (counts["n_acc"] * r_acc + counts["n_cost"] * r_cost)
    sample = ("one real oracle outer iteration on scans 0,1 at full size (%d pts each) from state '%s': %.1f s wall (search %.1f s, LM loop %.1f s: "
              "%d LM iterations, %d tries, %d correspondences; accumulate %.1f ns/corr/pass, cost %.1f ns/corr/pass); scaled to 8 scans: 56 directions over "
              "%d cores + C=%.3g x (%d accumulate + %d cost passes) as counted by the B200 run of the same step"
              % (args.scan_w * args.scan_h, args.state, dt, st["t_search"], st["t_inner"], st["inner_iterations"], st["lm_tries_total"], C_s,
                 r_acc * 1e9, r_cost * 1e9, cores, counts["C"], counts["n_acc"], counts["n_cost"]))
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample, "seconds_per_iteration_scaled": t_full,
            "sample_seconds": dt}


def run_reference(args):
    """--impl reference: ONE real outer iteration of the reference algorithm (oracle port: the reference itself needs PCL / FLANN / Eigen
    and cannot be built here) on the full 8 x 10M scans from the same state as the B200 arm. steps = 1 whatever --steps says: the serial
    accumulate / cost loops of PointToPlaneICPImpl::compute make one iteration take minutes."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    ensure_scans(range(args.scans), args.scan_w, args.scan_h)
    poses = state_poses(args)
    dt, st, cores = cpu_iteration(args, list(range(args.scans)), poses)
    v = 1.0 / dt
    cfg = workload(args)
    cfg.update({"correspondences": st["num_correspondences"], "inner_iterations": st["inner_iterations"], "lm_tries": st["lm_tries_total"],
                "lm_tries_per_iteration": st["tries"],
                "seconds": {"transform": st["t_transform"], "search": st["t_search"], "inner": st["t_inner"], "accumulate_passes": st["t_acc"],
                            "cost_passes": st["t_cost"]},
                "note": "one full outer iteration executed for real (not sampled, not scaled); --steps/--warmup are not repeated because one step takes minutes"})
    sample = "the full step: one real outer iteration on all %d scans (%d pts each), %d pair-directions, complete LM loop" % (
        args.scans, args.scan_w * args.scan_h, args.scans * (args.scans - 1))
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": 1, "warmup": 0,
           "ms_per_step": 1000.0 * dt, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 residuals/Jacobians, f64 accumulation", "data": "synthetic", "config": cfg,
           "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# --------------------------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------------------------
class _CudaArray:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.path = "/tmp/b2_clocks_%d_%d.csv" % (os.getpid(), index)
        self.proc = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            p = [x.strip() for x in line.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def hist(values):
    out = {}
    for v in values:
        out[str(int(v))] = out.get(str(int(v)), 0) + 1
    return out


def parity_check(args, g, poses, stride=89):
    """Bit-exact comparison of one pair-direction of the full-size run with the oracle's kd-tree search (FindCorrespondencesFast,
    icp_point_to_plane.cc:42-105) on every `stride`-th query: the oracle transforms scans 0 and 1 with the poses the search ran at,
    builds its kd-tree over ALL of scan 1 and answers the sampled queries of scan 0."""
    from oracle import oracle as orc
    orc.build()
    k = None
    for i, (s, t, c) in enumerate(g.pairs(with_lists=False)):
        if (s, t) == (0, 1):
            k = i
    if k is None:
        return {"checked": False, "why": "pair 0->1 not searched on this rank"}
    _, _, q, m, d2 = g.pair_correspondences(k)
    W, H = args.scan_w, args.scan_h
    a_xyz, a_nrm = load_scan(0, W, H); b_xyz, b_nrm = load_scan(1, W, H)
    ga, _ = orc.transform_cloud(a_xyz, a_nrm, poses[0]); gb, _ = orc.transform_cloud(b_xyz, b_nrm, poses[1])
    sel = np.arange(0, ga.shape[0], stride)
    t0 = time.perf_counter()
    qo, mo, do = orc.find_correspondences(ga[sel], gb, args.d, use_kdtree=True)
    t_or = time.perf_counter() - t0
    keep = (q % stride) == 0
    qg, mg, dg = q[keep] // stride, m[keep], d2[keep]
    ok = bool(np.array_equal(qg, qo) and np.array_equal(mg, mo) and np.array_equal(dg, do))
    return {"checked": True, "bit_exact": ok, "pair": "scan 0 -> scan 1", "queries_checked": int(sel.size), "matches_checked": int(qo.size),
            "gpu_matches_of_pair": int(q.size), "against": "oracle kd-tree over all %d target points (oracle/orc_icp.cc: find_correspondences)" % gb.shape[0],
            "oracle_seconds": t_or}


def run_b200(args):
    # stdout must carry exactly ONE JSON line: libraries that chat on fd 1 (NCCL prints its version banner there) are sent to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))
    from dataset_pipeline_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        if rank == 0:
            _lib.build()
        if world > 1:
            dist.barrier()
    import dataset_pipeline_b200 as b2

    W, H, NS = args.scan_w, args.scan_h, args.scans
    t_gen = time.perf_counter()
    if rank == 0:
        ensure_scans(range(NS), W, H)
    if world > 1:
        dist.barrier()
    clouds = []
    for i in range(NS):
        xyz, nrm = load_scan(i, W, H)
        px = torch.empty(xyz.shape, dtype=torch.float32, pin_memory=True); pn = torch.empty(nrm.shape, dtype=torch.float32, pin_memory=True)
        px.numpy()[:] = xyz; pn.numpy()[:] = nrm
        clouds.append((px, pn))
    t_gen = time.perf_counter() - t_gen
    start_poses, gt_poses = scene_poses(NS)
    base_poses = gt_poses if args.state == "gt" else start_poses
    npts = sum(c[0].shape[0] for c in clouds)
    stream = torch.cuda.Stream()          # a real (non-null) stream shared by torch events, NCCL ordering and the library
    torch.cuda.set_stream(stream)

    def allreduce(ptr, count, strm):
        # the ONE exchange of the data path: sum-allreduce of [H | b | cost | counters] (NCCL over NVLink), completed before returning
        t = torch.as_tensor(_CudaArray(ptr, count), device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        torch.cuda.current_stream().synchronize()

    comm = None
    if world > 1 and os.environ.get("B2_ALLREDUCE", "nccl") == "nccl":
        # library-owned NCCL communicator; the 128-byte id travels over torch.distributed
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(b2.Comm.unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        comm = b2.Comm(rank, world, bytes(idt.cpu().numpy().tobytes()), device=local)

    def make(poses):
        g = b2.PointToPlaneICP(device=local, rank=rank, world_size=world, allreduce=allreduce if (world > 1 and comm is None) else None,
                               stream=stream.cuda_stream, comm=comm, index_distance_hint=args.d, shard_uploads=comm is not None)
        g.upload_ms = []
        for (px, pn), T in zip(clouds, poses):
            t = time.perf_counter()
            g.AddPointCloud(px.numpy(), pn.numpy(), T)
            g.upload_ms.append(round(1e3 * (time.perf_counter() - t), 2))
        return g

    def set_poses(g, poses):
        for i, T in enumerate(poses):
            g.SetGlobalTCloud(i, T)

    def brief(st):
        return {"inner_iterations": st["inner_iterations"], "lm_tries": st["lm_tries_total"], "passes": st["passes"], "ms": round(st["ms_total"], 3)}

    g = make(start_poses)
    # ---- untimed: the whole alignment from the perturbed start (also the warm-up of allocations and of the static index) ----
    extra = None
    if not args.no_extra:
        traj = []
        conv = False
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for it in range(40):
            conv = g.Run(args.d, it, 1, THR, False)
            traj.append(brief(g.stats()))
            if conv:
                break
        torch.cuda.synchronize(); t_traj = time.perf_counter() - t0
        g.Run(args.d, len(traj), 1, THR, False)           # one more at the converged state
        extra = {"alignment_from_perturbed_start": {"outer_iterations_until_converged": len(traj), "converged": bool(conv), "seconds": t_traj,
                                                    "iterations_per_s": len(traj) / t_traj, "per_iteration": traj},
                 "converged_iteration": brief(g.stats())}
    for it in range(args.warmup):
        set_poses(g, base_poses)
        g.Run(args.d, it, 1, THR, False)
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    set_poses(g, base_poses)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step_stats, step_poses = [], []
    conv = False
    for it in range(args.steps):
        if args.state != "trajectory" or conv or it == 0:
            set_poses(g, base_poses)                       # host-side: eight 4x4 matrices
        step_poses.append([g.GetResultGlobalTCloud(i) for i in range(NS)])
        conv = g.Run(args.d, it, 1, THR, False)
        step_stats.append(g.stats())
    e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps / (ms / 1000.0)

    # ---- untimed A/B: the same step twice more with the optional overlap of K4 (pack) with K3 (search) switched on ----
    overlap_stats = None
    if not args.no_extra:
        g.set_option("pack_overlap", 1)
        for it in range(3):
            set_poses(g, base_poses)
            g.Run(args.d, it, 1, THR, False)
        overlap_stats = g.stats()
        g.set_option("pack_overlap", 0)

    # ---- parity of the timed path at full size: one pair-direction of the LAST timed step against the oracle ----
    parity = None
    if not args.no_parity and rank == 0:
        parity = parity_check(args, g, step_poses[-1])
    g.close()
    del g

    # ---- end to end through the public API with HOST buffers: create, 8 x AddPointCloud (H2D from pinned memory), Run, read poses ----
    e2e = None
    if not args.no_e2e:
        k_e2e = args.steps
        times, parts = [], None
        for s in range(1 + k_e2e):
            poses_s = step_poses[(s - 1) % len(step_poses)] if s else step_poses[0]
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ge = make(poses_s)
            t1 = time.perf_counter()
            ge.Run(args.d, 0, 1, THR, False)
            t2 = time.perf_counter()
            _ = [ge.GetResultGlobalTCloud(i) for i in range(NS)]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st_e = ge.stats()
            ge.close()
            parts = {"create_and_upload_ms": 1e3 * (t1 - t0), "add_cloud_ms": ge.upload_ms, "run_ms": 1e3 * (t2 - t1), "run_device_ms": st_e["ms_total"], "index_build_ms": st_e["ms_index_build"],
                     "destroy_ms": 1e3 * (time.perf_counter() - t0 - dt),
                     "run_phases_ms": {k: st_e["ms_" + k] for k in ("index", "search", "pack", "inner")}, "passes": st_e["passes"],
                     "searches_done_behind_uploads": st_e["searches_ahead"], "searches_in_run": st_e["search_launches"]}
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if s >= 1:
                times.append(dt)
        nv = 6 * (NS - 1)
        d2h = st_e["passes"] * (nv * nv + nv + 6) * 8 + NS * 4 * 148 * 6 * 4 + 4 * NS * (NS - 1) + 4 * NS
        h2d = npts * 24 if comm is None else sum(c[0].shape[0] * 24 for i, c in enumerate(clouds) if i % world == rank)   # rank 0's share when sharded
        e2e = {"value": len(times) / sum(times), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "step_ms": [round(1e3 * v, 2) for v in times],
               "steps": len(times), "last_step_breakdown": parts, "h2d_gb_per_s_in_upload": npts * 24 / 1e9 / max(parts["create_and_upload_ms"] * 1e-3, 1e-9),
               "note": "each step: b2_icp_create + 8 x b2_icp_add_cloud from pinned host memory (static index built, and on one GPU the pair-directions among the clouds "
                       "already resident searched, behind the uploads that follow) + b2_icp_run(1 iteration from the same poses as the corresponding timed step) + "
                       "b2_icp_get_pose + destroy; all of it inside the timed region"}

    # ---- the other half of BASELINE.json's metric: ImageRegistrator residual-evaluations/s (config 4; images sharded over the ranks) ----
    secondary = None
    if not args.no_secondary:
        try:
            import bench_reg
            torch.cuda.set_stream(torch.cuda.default_stream())
            secondary = bench_reg.secondary_line(world, rank, comm, local, cpu_baseline=(world == 1 and not args.no_cpu_baseline))
        except Exception as e:      # the ICP line must not be lost to the secondary benchmark
            secondary = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- rooflines (SURVEY.md §8d byte counts) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    src = "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    mean = lambda k: float(np.mean([s[k] for s in step_stats]))
    step_ms = ms / args.steps

    def traffic_of(name, algorithmic):
        tp = os.path.join(ROOT, "profiles", name)
        try:
            tj = json.load(open(tp))          # ncu capture (profiles/): DRAM bytes over algorithmic bytes of that capture, applied to this launch
            return tj["dram_over_algorithmic"] * algorithmic if "dram_over_algorithmic" in tj else tj.get("dram_bytes_per_launch")
        except Exception:
            return None

    acc_ms = mean("ms_accum_kernel_avg"); recs = mean("local_correspondences")
    achieved = (48.0 * recs / (acc_ms * 1e-3)) / 1e9 if acc_ms > 0 else 0.0
    roof_acc = {"bound": "hbm", "kernel": "k_accumulate_tma<WITH_H, NX> (K5, 48 B/correspondence/pass; a pass evaluates up to 4 LM tries on one read of the records)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None, "traffic": traffic_of("accumulate_traffic.json", 48.0 * recs),
                "peak_source": src, "algorithmic_bytes_per_launch": 48.0 * recs, "avg_launch_ms": acc_ms, "share_of_step": mean("passes") * acc_ms / step_ms}
    nn_ms = mean("ms_search_kernel_avg"); nn_launches = max(1.0, mean("search_launches"))
    nn_bytes = mean("search_algorithmic_bytes") / nn_launches
    overlapped = mean("packs_overlapped") > 0
    # with the pack overlapped the phase holds K3 AND K4: K4's algorithmic bytes per set are 8 Q (keys) + 116 Qm (two gathered 32 B rows,
    # the 4 B permutation entry, the 48 B record written)
    pack_bytes = (8.0 * (mean("search_algorithmic_bytes") - 8.0 * recs) / 24.0 + 116.0 * recs) / nn_launches if overlapped else 0.0
    nn_ach = ((nn_bytes + pack_bytes) / (nn_ms * 1e-3)) / 1e9 if nn_ms > 0 else 0.0
    roof_nn = {"bound": "hbm", "kernel": "k_nn_tiles (K3, 12Q+8Qm+12T B/pair-direction; instruction-issue / gather-latency bound in practice)"
                                          + (" with k_pack_tiles (K4, 8Q+116Qm B) of the previous sets running beside it" if overlapped else ""), "achieved": nn_ach,
               "peak": peak, "unit": "GB/s", "frac": nn_ach / peak if peak else None, "traffic": traffic_of("search_traffic.json", nn_bytes), "peak_source": src,
               "algorithmic_bytes_per_launch": nn_bytes + pack_bytes, "avg_launch_ms": nn_ms, "share_of_step": nn_launches * nn_ms / step_ms,
               "note": "launches of the pair-directions overlap on 4 streams: avg_launch_ms = search phase / launches; traffic = K3's share only"}
    if overlap_stats is not None:
        roof_nn["with_pack_overlapped"] = {"ms_search_and_pack": overlap_stats["ms_search"] + overlap_stats["ms_pack"], "ms_total": overlap_stats["ms_total"],
                                           "sets_packed_behind_searches": overlap_stats["packs_overlapped"],
                                           "note": "one untimed step with b2_icp_set_option(pack_overlap=1): K4 of a set runs while later sets are searched; "
                                                   "compare with ms_breakdown.search + ms_breakdown.pack of the timed steps"}
    # the dominant kernel of the step is the one the roofline key describes; the other is kept alongside
    roofline, roofline2 = (roof_nn, roof_acc) if roof_nn["share_of_step"] >= roof_acc["share_of_step"] else (roof_acc, roof_nn)
    cfg = workload(args)
    cfg.update({"parallelism": "pair-directions sharded over %d GPU(s), 1 allreduce of the normal equations per pass" % world,
                "l2": "inputs larger than L2 (%.1f GB of scans, %.1f GB of packed records per pass)" % (npts * 24 / 1e9, 48.0 * recs / 1e9),
                "correspondences": mean("num_correspondences"), "inner_iterations": mean("inner_iterations"), "lm_tries": mean("lm_tries_total"),
                "passes_per_step": mean("passes"), "sets_packed_behind_searches_per_step": mean("packs_overlapped"),
                "per_step": {"inner_iterations": hist(s["inner_iterations"] for s in step_stats), "passes": hist(s["passes"] for s in step_stats)},
                "lm": "reference semantics (tries evaluated in order); up to 4 tries ride on one streaming pass",
                "ms_breakdown": {"index": mean("ms_index"), "search": mean("ms_search"), "pack": mean("ms_pack"), "inner": mean("ms_inner")},
                "input_generation_s": t_gen})
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 residuals/Jacobians, f64 accumulation", "data": "synthetic", "config": cfg,
           "gpu_launches": int(sum(s["kernel_launches"] for s in step_stats)),
           "clocks": clocks, "roofline": roofline, "roofline_second_kernel": roofline2}
    if e2e:
        out["e2e"] = e2e
    if parity:
        out["parity_checked"] = parity
    if extra:
        out["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        counts = {"C": mean("num_correspondences"), "n_acc": int(round(mean("inner_iterations"))) + 1, "n_cost": int(round(mean("lm_tries_total")))}
        cb = cpu_baseline_sample(args, counts)
        out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if secondary is not None:
        out["secondary"] = secondary
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(out))
    sys.stdout.flush()
    os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
