#!/usr/bin/env python
"""Benchmark of the ICP hot path (BASELINE.json configs[1]): ICPScanAligner on 8 x 10M-point synthetic scans, point-to-plane.

    python bench.py --gpus N --steps K --warmup W            # this repo's B200 path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port) on the host cores

A step = one ICP outer iteration = one `icp.Run(d, it, 1, thr)` as issued by icp_scan_aligner.cc:343 (global-frame transform,
all pair-direction correspondence searches, the complete inner LM loop of PointToPlaneICPImpl::compute, pose composition).
Prints ONE JSON line (rank 0). Data: synthetic room scans (dataset_pipeline_b200/synth), inputs far larger than L2.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ICP iterations/sec (8x10M-pt scans, point-to-plane, d=0.01)"
UNIT = "iterations/s"
CACHE = os.environ.get("B2_SYNTH_CACHE", "/tmp/b2_synth_cache")


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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


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
    from dataset_pipeline_b200 import synth
    gts = []
    for i in range(nscans):
        T = np.eye(4); T[:3, :3] = synth.rot_xyz(0, 0, 0.35 * i); T[:3, 3] = synth.SCANNER_POSITIONS[i]
        gts.append(T)
    return synth.perturbed_poses(gts), gts


# --------------------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle port timed on a bounded sample, scaled to the metric's unit
# --------------------------------------------------------------------------------------------------------------------
def cpu_sample(args, counts=None):
    """Times the reference algorithm (oracle port, reference's parallel structure: OpenMP over pair-directions only, serial
    accumulate / cost loops) on a bounded sample of the workload and scales to one full outer iteration.

    Sample: scans 0 and 1 at FULL size. (a) global-frame transform of one full scan; (b) P concurrent single-thread probes,
    each = kd-tree build over a full 10M target + nearest-within-radius for every `stride`-th source point (P = host cores,
    as OpenMP would run P pair-directions at once); (c) the oracle ICP on the two scans with the source stride and the inner LM
    loop capped, for the per-correspondence accumulate / cost pass rates.
    Full iteration = 8 transforms + 56 (build + N queries) / min(P,56) + C * (n_acc * r_acc + n_cost * r_cost).
    C, n_acc, n_cost come from the B200 run of the same workload when given (`counts`), else from the sample itself
    (match fraction of the probes; LM counts of the capped run = a LOWER bound on the CPU time)."""
    from oracle import oracle as orc
    orc.build()
    W, H = args.scan_w, args.scan_h
    ensure_scans([0, 1], W, H)
    poses, _ = scene_poses(max(2, args.scans))
    a_xyz, a_nrm = load_scan(0, W, H); b_xyz, b_nrm = load_scan(1, W, H)
    cores = os.cpu_count() or 1
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    n_pts = a_xyz.shape[0]
    stride = max(1, n_pts // 400000)
    gb, _ = orc.transform_cloud(b_xyz, b_nrm, poses[1])     # first call also pays one-time costs: not timed
    t0 = time.perf_counter()
    ga, _ = orc.transform_cloud(a_xyz, a_nrm, poses[0])
    t_transform_pp = (time.perf_counter() - t0) / n_pts
    P = max(1, min(cores, 56))
    res = [None] * P

    def probe(k):
        src, tgt = (ga, gb) if k % 2 == 0 else (gb, ga)
        off = (k // 2) % stride
        res[k] = orc.time_search(src[off::stride], tgt, args.d)

    th = [threading.Thread(target=probe, args=(k,)) for k in range(P)]
    tw = time.perf_counter()
    [t.start() for t in th]; [t.join() for t in th]
    probe_wall = time.perf_counter() - tw
    nq = math.ceil(n_pts / stride)
    t_build = float(np.mean([r[0] for r in res])); t_q = float(np.mean([r[1] for r in res])) / nq
    frac = float(np.mean([r[2] for r in res])) / nq
    del ga, gb
    icp = orc.PointToPlaneICP(use_kdtree=True, inner_max_iterations=3)
    icp.set_query_stride(stride)
    icp.AddPointCloud(a_xyz, a_nrm, poses[0]); icp.AddPointCloud(b_xyz, b_nrm, poses[1])
    icp.Run(args.d, 0, 1, 1e-10, False)
    st = icp.stats()
    C_s = max(1, st["num_correspondences"])
    r_acc = st["t_acc"] / (C_s * max(1, st["inner_iterations"]))
    r_cost = st["t_cost"] / (C_s * max(1, st["lm_tries_total"]))
    ndirs = args.scans * (args.scans - 1)
    if counts:
        C, n_acc, n_cost, src = counts["C"], counts["n_acc"], counts["n_cost"], "B200 run"
    else:
        C, n_acc, n_cost, src = frac * ndirs * n_pts, st["inner_iterations"], st["lm_tries_total"], "sample (LM capped at 3: lower bound)"
    t_full = args.scans * n_pts * t_transform_pp + ndirs * (t_build + n_pts * t_q) / min(P, ndirs) + C * (n_acc * r_acc + n_cost * r_cost)
    sample = ("scans 0,1 at full size (%d pts): %d concurrent 1-thread probes (kd-tree build over the full target + every %d-th source "
              "point, d=%g) + oracle ICP on the pair (source stride %d, LM capped at 3) ; rates: transform %.1f ns/pt, build %.2f s, "
              "query %.2f us, accumulate %.1f ns/corr/pass, cost %.1f ns/corr/pass ; scaled with C=%.3g, n_acc=%d, n_cost=%d from %s"
              % (n_pts, P, stride, args.d, stride, t_transform_pp * 1e9, t_build, t_q * 1e6, r_acc * 1e9, r_cost * 1e9, C, n_acc, n_cost, src))
    return {"value": 1.0 / t_full, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
            "seconds_per_iteration": t_full, "probe_wall_s": probe_wall}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for s in range(args.warmup + args.steps):
        r = cpu_sample(args)
        if s >= args.warmup or args.warmup + args.steps <= 2:
            vals.append(r)
        if s == 0 and r["probe_wall_s"] > 60:      # keep the whole run within minutes on slow hosts
            vals = [r]
            break
    v = float(np.mean([x["value"] for x in vals]))
    last = vals[-1]
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
           "ms_per_step": 1000.0 / v, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 residuals, f64 accumulation",
           "data": "synthetic",
           "config": {"workload": "ICPScanAligner 8x10M-pt synthetic room scans, point-to-plane, d=0.01, 1 outer iteration per step",
                      "scans": args.scans, "points_per_scan": args.scan_w * args.scan_h, "note": "CPU oracle port of the reference algorithm, bounded sample scaled to one full iteration"},
           "cpu_baseline": {k: last[k] for k in ("unit", "cores", "kind", "sample")},
           "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    out["cpu_baseline"]["value"] = v
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
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
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
    poses, _ = scene_poses(NS)
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

    def make(start_poses=None):
        g = b2.PointToPlaneICP(device=local, rank=rank, world_size=world, allreduce=allreduce if (world > 1 and comm is None) else None,
                               stream=stream.cuda_stream, comm=comm)
        for (px, pn), T in zip(clouds, start_poses if start_poses is not None else poses):
            g.AddPointCloud(px.numpy(), pn.numpy(), T)
        return g

    thr = 1e-10
    g = make()
    for it in range(args.warmup):
        g.Run(args.d, it, 1, thr, False)
    sampler = ClockSampler(local) if rank == 0 else None
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    step_stats = []
    for it in range(args.warmup, args.warmup + args.steps):
        g.Run(args.d, it, 1, thr, False)
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
    final_poses = [g.GetResultGlobalTCloud(i) for i in range(NS)]
    g.close()
    del g

    # ---- end to end through the public API with HOST buffers: create, 8 x AddPointCloud (H2D from pinned memory), Run, read poses ----
    e2e = None
    if not args.no_e2e:
        k_e2e = max(1, min(args.steps, 3))
        times = []
        for s in range(1 + k_e2e):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ge = make(final_poses)          # the same kind of iteration as the timed ones: the alignment state after warm-up + steps
            t1 = time.perf_counter()
            ge.Run(args.d, args.warmup + args.steps, 1, thr, False)
            t2 = time.perf_counter()
            _ = [ge.GetResultGlobalTCloud(i) for i in range(NS)]
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            st_e = ge.stats()
            ge.close()
            parts = {"create_and_upload_ms": 1e3 * (t1 - t0), "run_ms": 1e3 * (t2 - t1), "run_device_ms": st_e["ms_total"], "destroy_ms": 1e3 * (time.perf_counter() - t0 - dt),
                     "run_phases_ms": {k: st_e["ms_" + k] for k in ("index", "search", "pack", "inner")}, "passes": st_e["passes"]}
            if world > 1:
                t = torch.tensor([dt], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            if s >= 1:
                times.append(dt)
        nv = 6 * (NS - 1)
        d2h = st_e["passes"] * (nv * nv + nv + 6) * 8 + NS * 2 * 148 * 6 * 4 + 8 * NS * (NS - 1) + 4 * NS
        e2e = {"value": len(times) / sum(times), "unit": UNIT, "h2d_bytes_per_step": int(npts * 24), "d2h_bytes_per_step": int(d2h),
               "steps": len(times), "last_step_breakdown": parts, "h2d_gb_per_s_in_upload": npts * 24 / 1e9 / max(parts["create_and_upload_ms"] * 1e-3, 1e-9), "note": "each step: b2_icp_create + 8 x b2_icp_add_cloud from pinned host memory (poses = the state the timed iterations ended in) + b2_icp_run(1 iteration) + b2_icp_get_pose + destroy"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (k_accumulate: 48 algorithmic bytes per correspondence per pass) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    acc_ms = float(np.mean([s["ms_accum_kernel_avg"] for s in step_stats]))
    recs = float(np.mean([s["local_correspondences"] for s in step_stats]))
    achieved = (48.0 * recs / (acc_ms * 1e-3)) / 1e9 if acc_ms > 0 else 0.0
    traffic = None
    tp = os.path.join(ROOT, "profiles", "accumulate_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))      # ncu capture (profiles/): DRAM bytes over algorithmic bytes of that capture, applied to this launch
            traffic = tj["dram_over_algorithmic"] * 48.0 * recs if "dram_over_algorithmic" in tj else tj.get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    src = "MEASURED_PEAKS.json (of measured)" if peaks else "fallback 6650 GB/s (of fallback)"
    mean = lambda k: float(np.mean([s[k] for s in step_stats]))
    roof_acc = {"bound": "hbm", "kernel": "k_accumulate_tma<WITH_H, NX> (K5, 48 B/correspondence/pass; a pass evaluates up to 4 LM tries on one read of the records)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": src,
                "algorithmic_bytes_per_launch": 48.0 * recs, "avg_launch_ms": acc_ms,
                "share_of_step": mean("passes") * acc_ms / (ms / args.steps)}
    nn_ms = mean("ms_search_kernel_avg"); nn_launches = max(1.0, mean("search_launches"))
    nn_bytes = mean("search_algorithmic_bytes") / nn_launches
    nn_ach = (nn_bytes / (nn_ms * 1e-3)) / 1e9 if nn_ms > 0 else 0.0
    nn_traffic = None
    tp = os.path.join(ROOT, "profiles", "search_traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            nn_traffic = tj["dram_over_algorithmic"] * nn_bytes if "dram_over_algorithmic" in tj else tj.get("dram_bytes_per_launch")
        except Exception:
            nn_traffic = None
    roof_nn = {"bound": "hbm", "kernel": "k_nn_radius1 (K3, 12Q+8Qm+12T B/pair-direction; gather/L2-latency bound in practice)", "achieved": nn_ach,
               "peak": peak, "unit": "GB/s", "frac": nn_ach / peak if peak else None, "traffic": nn_traffic, "peak_source": src,
               "algorithmic_bytes_per_launch": nn_bytes, "avg_launch_ms": nn_ms, "share_of_step": nn_launches * nn_ms / (ms / args.steps)}
    # the dominant kernel of the step is the one the roofline key describes; the other is kept alongside
    roofline, roofline2 = (roof_nn, roof_acc) if roof_nn["share_of_step"] >= roof_acc["share_of_step"] else (roof_acc, roof_nn)
    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
           "dtype": "f32 residuals/Jacobians, f64 accumulation", "data": "synthetic",
           "config": {"workload": "ICPScanAligner 8x10M-pt synthetic room scans, point-to-plane, d=0.01, 1 outer iteration per step",
                      "scans": NS, "points_per_scan": int(npts // NS), "parallelism": "pair-directions sharded over %d GPU(s), 1 allreduce of the normal equations per pass" % world,
                      "l2": "inputs larger than L2 (%.1f GB of scans, %.1f GB of packed records per pass)" % (npts * 24 / 1e9, 48.0 * recs / 1e9),
                      "correspondences": mean("num_correspondences"), "inner_iterations": mean("inner_iterations"), "lm_tries": mean("lm_tries_total"),
                      "passes_per_step": mean("passes"), "lm": "reference semantics (tries evaluated in order); up to 4 tries ride on one streaming pass",
                      "ms_breakdown": {"index": mean("ms_index"), "search": mean("ms_search"), "pack": mean("ms_pack"), "inner": mean("ms_inner")},
                      "input_generation_s": t_gen},
           "gpu_launches": int(sum(s["kernel_launches"] for s in step_stats)),
           "clocks": clocks, "roofline": roofline, "roofline_second_kernel": roofline2}
    if e2e:
        out["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        counts = {"C": mean("num_correspondences"), "n_acc": int(round(mean("inner_iterations"))), "n_cost": int(round(mean("lm_tries_total")))}
        cb = cpu_sample(args, counts)
        out["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
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
