#!/usr/bin/env python
"""bench.py — scans/sec of the per-scan registration hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1]): synthetic K64 scans (64 x 2048 rays, ~129 k returns)
rotating in place (yaw += 3 deg/scan) at poses sampled over a frozen local map of exactly 2^20 occupied
voxels (voxel 0.5 m, <= 20 points/voxel, culling off) built by streaming the T00 trajectory; initial-pose error
U[+-0.3 m, +-1 deg]; lidar3d-default ICP (pt2pt matcher, GN x2, Geman-McClure, 300 iterations cap).
One "step" = one batch of B scans through filter_1st_pass -> ICP align (no map insert: the map is frozen).

  value     scans/s, inputs resident in HBM when the timed region starts (mlo_scan_register_batch_resident)
  e2e       scans/s through the C ABI with pinned HOST buffers (H2D + D2H inside the timed region)
  roofline  fused NN+residual kernel: algorithmic bytes (SURVEY.md §8(d) formula) / CUDA-event time
  cpu_baseline  the CPU oracle (our restatement of mp2p_icp/mola_metric_maps — NOT the upstream binary) timed on
                this box's host cores on a bounded sample of the same scans
  --impl reference  times only that CPU path (the reference's own binary cannot be built: DESIGN.md)

  --workload sequence|ndt  (not the bench line) BASELINE configs[2]/[3]/[4]: whole sequences through the C++ host
                orchestrator; the --sequences S sequences of a GPU advance in lock step as one fleet (mlo_fleet_*).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from mola_lidar_odometry_b200 import capi, synth  # noqa: E402

TARGET_VOXELS = 1 << 20
MAP_VOXEL = 0.5
MAP_CAP = 20
EST_RANGE = 100.0
SIGMA = 2.0
WORKLOAD = "config[1]: K64 64x2048 rotating scans vs frozen 2^20-voxel map (0.5 m, cap 20), lidar3d-default ICP"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


# stdout carries exactly ONE JSON line: native libraries (e.g. "NCCL version ..." banners) also write to fd 1, so the
# real stdout is kept aside and fd 1 is pointed at stderr for everything else.
_REAL_STDOUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


# ---------------------------------------------------------------------------------------------- workload
class Workload:
    def __init__(self, n_query: int, seed: int, threads: int):
        self.scene = synth.Scene(42)
        self.traj = synth.trajectory_T00(8000, seed=7)  # T00-shaped drive, long enough to cover 2^20 voxels
        self.T0 = self.traj[0]
        self.fp = capi.filter1_default(EST_RANGE)
        self.threads = threads
        self.n_query = n_query
        self.seed = seed

    def map_poses(self, stride=2):
        return list(range(0, len(self.traj), stride))

    def gen_scans(self, idxs, seed0=1000):
        with ThreadPoolExecutor(self.threads) as ex:
            return list(ex.map(lambda k: self.scene.scan(self.traj[k], scan_seed=seed0 + k), idxs))

    def rel(self, k):
        return synth.relative(self.T0, self.traj[k])

    def query_set(self, max_pose_idx: int):
        """n_query scans: pose k sampled uniformly over the mapped part of the trajectory, yaw rotated by 3 deg x i."""
        rng = np.random.default_rng(self.seed)
        ks = rng.integers(0, max_pose_idx, self.n_query)
        poses_w, gts, inits = [], [], []
        for i, k in enumerate(ks):
            Tw = synth.compose(self.traj[k], synth.pose34(0, 0, 0, np.deg2rad(3.0 * i)))
            poses_w.append(Tw)
            gt = synth.relative(self.T0, Tw)
            gts.append(gt)
            inits.append(synth.perturb(gt, rng, 0.3, 1.0))
        with ThreadPoolExecutor(self.threads) as ex:
            scans = list(ex.map(lambda a: self.scene.scan(a[1], scan_seed=500000 + self.seed * 1000 + a[0]),
                                enumerate(poses_w)))
        return scans, np.stack(gts), np.stack(inits)


def stream_map(wl: Workload, filt, insert, stats, tag: str):
    """Stream T00 scans (every 2nd pose) through filter + insert until exactly 2^20 voxels exist.
    Returns the construction recipe [(pose index, points used)] so the other side can replay it."""
    idxs = wl.map_poses(2)
    t0 = time.time()
    layers = []
    for c0 in range(0, len(idxs), 64):
        ks = idxs[c0:c0 + 64]
        scans = wl.gen_scans(ks)
        layers_a = filt(scans)
        for k, a in zip(ks, layers_a):
            nv = stats()[0]
            pose = wl.rel(k)
            if nv + len(a) < TARGET_VOXELS:
                insert(a, pose)
                layers.append((k, len(a)))
                continue
            # close to the target: feed points in small chunks so the map lands on exactly 2^20 voxels
            i = 0
            while i < len(a) and nv < TARGET_VOXELS:
                step = 1 if TARGET_VOXELS - nv <= 256 else 256
                insert(a[i:i + step], pose)
                i += step
                nv = stats()[0]
            layers.append((k, i))
            if nv >= TARGET_VOXELS:
                log(f"[bench] {tag} map frozen: {stats()} from {len(layers)} scans (last pose {k}) in {time.time() - t0:.1f}s")
                return layers
    raise RuntimeError("trajectory exhausted before reaching 2^20 voxels")


def build_gpu_map(ctx, wl: Workload, capacity: int, rank: int = 0, world: int = 1):
    """Every rank ends with the same map.  With several ranks the scan synthesis + device filtering of each chunk
    of 64 poses is split across ranks and the decimated layers are all-gathered over NCCL (set-up only: the timed
    region has no collective)."""
    from mola_lidar_odometry_b200.api import LocalMap
    gmap = LocalMap(ctx, MAP_VOXEL, MAP_CAP, 0.0, capacity)
    if world == 1:
        layers = stream_map(wl, lambda scans: [ctx.filter_1st_pass(r, wl.fp)[0] for r in scans], gmap.insert, gmap.stats,
                            "device")
        return gmap, layers
    import torch
    import torch.distributed as dist
    MAXP = 16384
    gen_all = wl.gen_scans

    def gen_share(idxs):       # stream_map asks for the scans of a chunk: produce only this rank's share
        mine = idxs[rank::world]
        return [(i, s) for i, s in zip(mine, gen_all(mine))] if mine else []

    def filt_shared(tagged):
        n_chunk = filt_shared.chunk_len
        per = (n_chunk + world - 1) // world
        buf = torch.zeros((per, MAXP, 3), dtype=torch.float32, device="cuda")
        cnt = torch.zeros((per,), dtype=torch.int32, device="cuda")
        for j, (_, raw) in enumerate(tagged):
            a = ctx.filter_1st_pass(raw, wl.fp)[0]
            assert len(a) <= MAXP
            buf[j, :len(a)] = torch.from_numpy(a).cuda()
            cnt[j] = len(a)
        bufs = [torch.empty_like(buf) for _ in range(world)]
        cnts = [torch.empty_like(cnt) for _ in range(world)]
        dist.all_gather(bufs, buf)
        dist.all_gather(cnts, cnt)
        out = []
        for i in range(n_chunk):                      # chunk position i was produced by rank i % world as its item i // world
            r, j = i % world, i // world
            n = int(cnts[r][j])
            out.append(bufs[r][j, :n].cpu().numpy())
        return out

    class _Shim:
        """Adapts stream_map's (gen, filt) protocol: gen returns a tagged share, filt gathers the whole chunk."""
    orig_gen = wl.gen_scans

    def gen_hook(idxs, seed0=1000):
        filt_shared.chunk_len = len(idxs)
        return gen_share(list(idxs))
    wl.gen_scans = gen_hook
    try:
        layers = stream_map(wl, filt_shared, gmap.insert, gmap.stats, f"device[rank {rank}]")
    finally:
        wl.gen_scans = orig_gen
    return gmap, layers


def _oracle_filter(wl):
    from oracle import oracle_py as O

    def f(scans):
        with ThreadPoolExecutor(wl.threads) as ex:
            return list(ex.map(lambda r: O.filter_1st_pass(r, wl.fp)[0], scans))
    return f


def build_oracle_map(wl: Workload, layers):
    """The same map on the CPU side (oracle's own filter + insert), replaying the recipe, for the CPU baseline."""
    from oracle import oracle_py as O
    omap = O.OracleMap(MAP_VOXEL, MAP_CAP, 0.0)
    t0 = time.time()
    filt = _oracle_filter(wl)
    for c0 in range(0, len(layers), 64):
        part = layers[c0:c0 + 64]
        for (k, n), a in zip(part, filt(wl.gen_scans([k for k, _ in part]))):
            omap.insert(a[:n], wl.rel(k))
    log(f"[bench] oracle map: {omap.stats()} in {time.time() - t0:.1f}s")
    return omap


def plan_map_layers(wl: Workload):
    """Reference arm (no GPU): the same construction driven by the oracle."""
    from oracle import oracle_py as O
    omap = O.OracleMap(MAP_VOXEL, MAP_CAP, 0.0)
    layers = stream_map(wl, _oracle_filter(wl), omap.insert, omap.stats, "oracle")
    return omap, layers


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.p = None
        self.lines = []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_scans_per_sec(omap, scans, inits, fp, n_threads: int, budget_s: float, max_scans: int, warm_scans: int = 0,
                      order=None):
    """filter_1st_pass + align with the CPU oracle: ONE persistent pool of `n_threads` workers pulling scans off one queue
    (process-per-sequence is the reference's own parallelism, eval/cli_kitti.sh:23) - no per-step join, thread start-up
    outside the timed region, `warm_scans` untimed scans first (>= 3).  Bounded by `budget_s` /
    `max_scans`.  Returns (scans/s, scans done, seconds, mean per-bucket ms, median-of-5-segments scans/s)."""
    from oracle import oracle_py as O
    ips = [capi.IcpParamsOwner(sigma=SIGMA) for _ in range(n_threads)]
    order = list(range(len(scans))) if order is None else list(order)
    todo = order[:max_scans]
    lock = threading.Lock()
    state = {"next": 0, "t0": None, "go": False}
    done = []
    ready = threading.Barrier(n_threads + 1)

    def worker(tid):
        for w in range(warm_scans // n_threads + (1 if tid < warm_scans % n_threads else 0)):   # untimed warm-up
            i = todo[(tid + w * n_threads) % len(todo)]
            O.scan_register(omap, scans[i], fp, inits[i], ips[tid].p)
        ready.wait()      # every worker is warm: the main thread starts the clock
        ready.wait()
        while True:
            with lock:
                k = state["next"]
                if k >= len(todo) or time.perf_counter() - state["t0"] > budget_s:
                    return
                state["next"] += 1
            i = todo[k]
            res, ms = O.scan_register(omap, scans[i], fp, inits[i], ips[tid].p)
            t = time.perf_counter()
            with lock:
                done.append((t, i, res.n_iterations, ms))

    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    for t in th:
        t.start()
    ready.wait()
    state["t0"] = time.perf_counter()
    ready.wait()
    for t in th:
        t.join()
    t_end = max([d[0] for d in done], default=state["t0"])
    dt = t_end - state["t0"]
    ms = np.array([d[3] for d in done])
    # median of 5 consecutive segments of the run (by completion time)
    seg = None
    if len(done) >= 10:
        ts = np.sort(np.array([d[0] for d in done])) - state["t0"]
        edges = [0.0] + [float(ts[(j + 1) * len(ts) // 5 - 1]) for j in range(5)]
        counts = [(j + 1) * len(ts) // 5 - j * len(ts) // 5 for j in range(5)]
        seg = float(np.median([c / max(e1 - e0, 1e-9) for c, e0, e1 in zip(counts, edges[:-1], edges[1:])]))
    return len(done) / max(dt, 1e-9), len(done), dt, (ms.mean(0) if len(done) else np.zeros(3)), seg


def config_dict(B, scans, map_stats, world):
    """The `config` of the bench line: identical for the GPU arm and the --impl reference arm."""
    return {"workload": WORKLOAD, "scans_per_step_per_gpu": B, "points_per_scan": int(np.mean([len(s) for s in scans])),
            "map_voxels": int(map_stats[0]), "map_points": int(map_stats[1]),
            "l2_policy": "inputs larger than L2: 4 rotating windows of B raw scans + 400 MB map working set",
            "parallelism": f"replicas x{world} (one rank per GPU, map replicated, scans sharded)"}


N_WINDOWS = 4


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU path alone (oracle port: the upstream binary cannot be built here)
    on the SAME workload, config and step size as the GPU arm: steps x B scans stream through one persistent pool of all
    host cores (no per-step join); K steps are timed after W warm-up steps' worth of scans (capped at 4 per core)."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = args.batch
    wl = Workload(B * N_WINDOWS, seed=args.seed, threads=cores)
    omap, layers = plan_map_layers(wl)
    scans, gts, inits = wl.query_set(layers[-1][0])
    order = [(s * B + j) % len(scans) for s in range(args.steps) for j in range(B)]
    warm = min(args.warmup * B, 4 * cores)
    value, n, dt, _, seg = cpu_scans_per_sec(omap, scans, inits, wl.fp, cores, 1e9, len(order), warm_scans=max(3, warm), order=order)
    line = {"metric": "scans/sec", "value": value, "unit": "scans/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
            "impl": "reference",
            "config": config_dict(B, scans, omap.stats(), args.gpus),
            "cpu_baseline": {"value": value, "unit": "scans/s", "cores": cores, "kind": "port",
                             "sample": f"{n} scans = {args.steps} steps x {B}, one persistent pool of {cores} workers "
                                       f"(independent scans in flight, like eval/cli_kitti.sh), {max(3, warm)} warm-up scans",
                             "median_of_5_segments": seg,
                             "note": "CPU = our restatement of mp2p_icp/mola_metric_maps, not the upstream binary (unbuildable here)"},
            "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ---------------------------------------------------------------------------------------------- sequence workloads
SEQ_WORKLOADS = {
    # BASELINE.json configs[2]/[4]: T00-shaped drives x K64, lidar3d-default.yaml (pt2pt matcher, GN + Geman-McClure)
    "sequence": dict(yaml="lidar3d-default.yaml", sensor="K64",
                     label="config[2]/[4]: T00 x K64 sequences, lidar3d-default.yaml, full odometry loop"),
    # configs[2] point-to-plane variant: the same drives through the NDT pipeline (Matcher_Point2Plane first)
    "sequence_pt2pl": dict(yaml="lidar3d-ndt.yaml", sensor="K64",
                           label="config[2] point-to-plane: T00 x K64 sequences, lidar3d-ndt.yaml (Matcher_Point2Plane + NDT map)"),
    # configs[3]: 128-beam 230 k-point sweeps, lidar3d-ndt.yaml
    "ndt": dict(yaml="lidar3d-ndt.yaml", sensor="O128", label="config[3]: T00 x O128 (230 k rays), lidar3d-ndt.yaml"),
}


def run_fleet_workload(ctx, kind: str, S: int, N: int, rank: int, world: int, cores: int, with_cpu: bool, cpu_scans: int,
                       chunk_steps: int = 0, prefetch: bool = True, gpu_arm: bool = True):
    """S independent T00-shaped sequences of N scans through the FULL odometry loop of the C++ host layer (filter ->
    motion-model guess + prior -> ICP -> quality gate -> adaptive sigma -> keyframe -> map insert + cull) with the
    reference's pipeline YAML.  GPU arm: the S sequences advance in lock step as one fleet (mlo_fleet_*): pinned host
    scans in (x, y, z packed), host results out, the next step's upload overlapped with the ICP.  CPU arm: the SAME
    orchestrator over the oracle backend, one thread per sequence (the reference's process-per-sequence,
    eval/cli_kitti.sh:23), on min(S, cores) sequences and their first `cpu_scans` scans.  Scans are ray-cast on the GPU
    (synth/synth_gpu.cu, untimed) chunk by chunk; only lock steps are timed.  Returns the sub-record."""
    import torch
    from mola_lidar_odometry_b200.synth.gpu import GpuSynth
    from oracle import oracle_py as O
    w = SEQ_WORKLOADS[kind]
    yaml_path = ROOT / "pipelines" / w["yaml"]
    sensor = getattr(synth, w["sensor"])
    os.environ.setdefault("MOLA_OPTIMIZE_TWIST", "false")   # benchmark settings of SURVEY.md §8(d): deskew is row f1
    os.environ.setdefault("MOLA_INITIAL_VX", "8.0")
    dev = torch.device("cuda", torch.cuda.current_device())
    scene = synth.Scene(42)
    gs = GpuSynth(scene, dev)
    seeds = [7 + rank * S + i for i in range(S)]
    trajs = [synth.trajectory_T00(N + 5, seed=sd) for sd in seeds]
    n_rays = sensor.n_beams * sensor.n_az
    if chunk_steps <= 0:   # ~2.5 GB of pinned host memory per chunk
        chunk_steps = max(1, min(N, int(2.5e9 / (S * n_rays * 12))))
    pinned = torch.empty((S * chunk_steps * n_rays, 3), dtype=torch.float32).pin_memory()
    pinned_np = pinned.numpy()
    fleet = None
    if gpu_arm:
        from mola_lidar_odometry_b200.host_api import LidarOdometryFleet
        fleet = LidarOdometryFleet(ctx, yaml_path, S)
    n_cpu = min(S, cores) if with_cpu else 0
    cpu_lo = [O.OracleLidarOdometry(yaml_path) for _ in range(n_cpu)]
    gpu_wall, cpu_wall, cpu_done, launches0 = 0.0, 0.0, 0, (ctx.launch_count if ctx else 0)
    gpu_poses = [[] for _ in range(S)]
    cpu_poses = [[] for _ in range(n_cpu)]
    its, pts_sum, h2d = 0, 0, 0
    prior_scans = 0
    for k0 in range(0, N, chunk_steps):
        k1 = min(N, k0 + chunk_steps)
        # ---- synthesis of this chunk (untimed): step-major order, scan (k, s) at index (k - k0) * S + s
        poses = np.stack([trajs[s][k] for k in range(k0, k1) for s in range(S)])
        sds = np.array([seeds[s] * 100000 + k for k in range(k0, k1) for s in range(S)], dtype=np.uint64)
        flat, offs = gs.scan_batch(poses, sds, sensor)
        offs_h = offs.cpu().numpy()
        pinned[:offs_h[-1]].copy_(flat[:, :3])
        torch.cuda.synchronize()
        del flat
        views = [[pinned_np[offs_h[(k - k0) * S + s]:offs_h[(k - k0) * S + s + 1]] for s in range(S)] for k in range(k0, k1)]
        pts_sum += int(offs_h[-1])
        # ---- GPU arm: lock steps, timed
        if fleet is not None:
            t0 = time.perf_counter()
            for k in range(k0, k1):
                if prefetch and k + 1 < k1:
                    fleet.prefetch(views[k + 1 - k0])
                outs = fleet.on_lidar(views[k - k0], [0.1 * k] * S, as_arrays=True)
                for s in range(S):
                    gpu_poses[s].append(outs["pose_3x4"][s].copy())
                its += int(outs["icp_iterations"].sum())
                prior_scans += int(outs["icp_had_prior"].sum())
            gpu_wall += time.perf_counter() - t0
            h2d += int(offs_h[-1]) * 12
        # ---- CPU arm: one thread per sequence over the same clouds
        if n_cpu and k0 < cpu_scans:
            kk1 = min(k1, cpu_scans)

            def drive(i):
                for k in range(k0, kk1):
                    o = cpu_lo[i].on_lidar(views[k - k0][i], 0.1 * k)
                    cpu_poses[i].append(o.pose.copy())
            t0 = time.perf_counter()
            with ThreadPoolExecutor(n_cpu) as ex:
                list(ex.map(drive, range(n_cpu)))
            cpu_wall += time.perf_counter() - t0
            cpu_done += n_cpu * (kk1 - k0)
    rec = {"workload": w["label"], "sequences_per_gpu": S, "scans_per_sequence": N, "points_per_scan": pts_sum // max(1, S * N),
           "host_layout": "x,y,z float32 packed (12 B/pt), pinned; scans ray-cast on the GPU beforehand (untimed)"}
    if fleet is not None:
        rec.update({"value": S * N / gpu_wall, "unit": "scans/s", "ms_per_lock_step": 1e3 * gpu_wall / N, "wall_s": gpu_wall,
                    "mean_icp_iterations": its / max(1, S * (N - 1)), "scans_with_prior": prior_scans,
                    "h2d_bytes_per_step": h2d // N, "gpu_launches": int(ctx.launch_count - launches0),
                    "phases_ms_per_step": fleet.phase_times()})
        fleet.close()
    if n_cpu:
        rec["cpu_baseline"] = {"value": cpu_done / cpu_wall, "unit": "scans/s", "cores": n_cpu, "kind": "port",
                               "sample": f"{n_cpu} sequences x first {min(N, cpu_scans)} scans, one thread per sequence "
                                         f"(same C++ orchestrator over the oracle backend)"}
        if fleet is not None:
            # Trajectory parity.  A sequence is a closed loop: after the first scan whose result differs beyond rounding
            # (different summation order -> a pairing or a stall test on the edge flips once in ~1e8 query-iterations) the
            # two arms no longer see the same map and initial guess, and both drift within the algorithm's own convergence
            # tolerance.  Per-scan parity is therefore asserted AT that first deviation, where the inputs were still
            # identical to ~1e-12: <= 1 mm / 0.01 deg (north_star).  The whole-trajectory figures are reported.
            dt, dr, ape, ape_cpu, first_dev, identical = [], [], [], [], [], 0
            for i in range(n_cpu):
                seq_dt = []
                for k, cp in enumerate(cpu_poses[i]):
                    e = O.pose_error(gpu_poses[i][k], cp)
                    seq_dt.append(e)
                dt += [e[0] for e in seq_dt]
                dr += [e[1] for e in seq_dt]
                dev = next((k for k, e in enumerate(seq_dt) if e[0] > 1e-9 or e[1] > 1e-8), None)
                if dev is None:
                    identical += 1
                else:
                    first_dev.append({"sequence": i, "scan": dev, "trans_m": seq_dt[dev][0], "rot_deg": seq_dt[dev][1]})
                gp, cpp = np.stack(gpu_poses[i][:len(cpu_poses[i])]), np.stack(cpu_poses[i])
                gt = np.stack([synth.relative(trajs[i][0], trajs[i][k]) for k in range(len(gp))])
                ape.append(float(np.sqrt(np.mean(np.sum((gp[:, :, 3] - gt[:, :, 3]) ** 2, axis=1)))))
                ape_cpu.append(float(np.sqrt(np.mean(np.sum((cpp[:, :, 3] - gt[:, :, 3]) ** 2, axis=1)))))
            rec["parity_vs_oracle"] = {
                "scans": len(dt), "sequences": n_cpu, "sequences_identical_to_1e-9_m": identical,
                "first_deviations": first_dev[:8],
                "max_trans_m_at_first_deviation": max([d["trans_m"] for d in first_dev], default=0.0),
                "max_rot_deg_at_first_deviation": max([d["rot_deg"] for d in first_dev], default=0.0),
                "max_trans_m": float(max(dt)), "max_rot_deg": float(max(dr)), "ape_rmse_m": float(np.sqrt(np.mean(np.square(dt)))),
                "tolerance": "1e-3 m / 1e-2 deg per scan at the first deviation of each sequence (identical inputs up to there), asserted"}
            rec["ape_rmse_vs_ground_truth_m"] = {"gpu": float(np.mean(ape)), "cpu_oracle": float(np.mean(ape_cpu))}
            rec["speedup_vs_cpu"] = rec["value"] / rec["cpu_baseline"]["value"]
            pv = rec["parity_vs_oracle"]
            assert pv["max_trans_m_at_first_deviation"] <= 1e-3 and pv["max_rot_deg_at_first_deviation"] <= 1e-2, pv
    return rec


def run_sequences(args, rank, local_rank, world):
    """`--workload sequence|sequence_pt2pl|ndt`: one whole-sequence workload as its own JSON line (see run_fleet_workload)."""
    cores = os.cpu_count() or 1
    import torch
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank)
    S, N = args.sequences, args.scans
    line = {"metric": "scans/sec", "unit": "scans/s", "n_gpus": world, "steps": N, "warmup": 0, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
            "config": {"workload": SEQ_WORKLOADS[args.workload]["label"], "sequences_per_gpu": S, "scans_per_sequence": N}}
    if args.impl == "reference":
        if rank != 0:
            return
        rec = run_fleet_workload(None, args.workload, S, N, 0, 1, cores, True, N, gpu_arm=False)
        v = rec["cpu_baseline"]["value"]
        line.update({"impl": "reference", "value": v, "ms_per_step": 1e3 * S / max(v, 1e-9), "cpu_baseline": rec["cpu_baseline"],
                     "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0})
        emit(line)
        return
    import torch.distributed as dist
    from mola_lidar_odometry_b200.api import Context
    from mola_lidar_odometry_b200.host_api import ScanOutput
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Context(local_rank)
    run_fleet_workload(ctx, args.workload, S, min(N, 12), rank, world, cores, False, 0, prefetch=not args.no_prefetch)   # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    rec = run_fleet_workload(ctx, args.workload, S, N, rank, world, cores, rank == 0 and not args.no_cpu_baseline,
                             min(N, args.cpu_scans), prefetch=not args.no_prefetch)
    t = torch.tensor([rec["wall_s"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    value = world * S * N / float(t[0])
    if rank == 0:
        line.update({"value": value, "ms_per_step": 1e3 * float(t[0]) / N,
                     "e2e": {"value": value, "unit": "scans/s", "h2d_bytes_per_step": rec["h2d_bytes_per_step"],
                             "d2h_bytes_per_step": int(S * C.sizeof(ScanOutput)),
                             "note": "pinned host scans in, host results out on every lock step (mlo_fleet_on_lidar)", "numa": numa},
                     "gpu_launches": rec["gpu_launches"], "cpu_baseline": rec.get("cpu_baseline"),
                     "quality": {"parity_vs_oracle": rec.get("parity_vs_oracle"), "ape_rmse_vs_gt_m": rec.get("ape_rmse_vs_ground_truth_m"),
                                 "mean_icp_iterations": rec["mean_icp_iterations"], "scans_with_prior": rec["scans_with_prior"]},
                     "phases": {"host_wall_timed_pass": rec["phases_ms_per_step"], "device_events_pass": {}}, "roofline": None})
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


def bind_near_gpu(idx: int):
    """Run this rank on the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) so that the pinned
    staging buffers it allocates are NUMA-local to the GPU's root port.  Returns a short description or None."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(idx)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        base = Path("/sys/bus/pci/devices") / bdf
        cpus = set()
        for part in (base / "local_cpulist").read_text().strip().split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        node = (base / "numa_node").read_text().strip()
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "numa_node": int(node), "cpus": len(cpus)}
    except Exception as e:  # sysfs layout differs / not permitted: run unbound
        return {"error": str(e)[:80]}


# ---------------------------------------------------------------------------------------------- GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="scans per step per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sub-sequences", type=int, default=32, help="sequences per GPU of the configs[4] sub-record")
    ap.add_argument("--sub-scans", type=int, default=1000, help="scans per sequence of the configs[4] sub-record")
    ap.add_argument("--cpu-scans", type=int, default=1000, help="sequence workloads: scans per sequence given to the CPU arm")
    ap.add_argument("--sub-records", default="auto", choices=["auto", "all", "fleet", "none"],
                    help="config1 line: also run BASELINE configs[2]/[3]/[4] and attach them as sub-records "
                         "(auto = all on one GPU, the fleet of configs[4] on several)")
    ap.add_argument("--workload", default="config1", choices=["config1", "sequence", "sequence_pt2pl", "ndt"],
                    help="config1 = the headline (default); sequence = BASELINE configs[2]/[4] full odometry loop; "
                         "ndt = configs[3] (lidar3d-ndt.yaml, O128 sensor)")
    ap.add_argument("--sequences", type=int, default=32, help="independent sequences per GPU (sequence/ndt workloads)")
    ap.add_argument("--no-prefetch", action="store_true", help="sequence workloads: no overlapped upload of the next step")
    ap.add_argument("--scans", type=int, default=120, help="scans per sequence (sequence/ndt workloads)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.workload != "config1":
        run_sequences(args, rank, local_rank, world)
        return
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from mola_lidar_odometry_b200.api import Context

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = bind_near_gpu(local_rank)   # before any pinned allocation: first touch places the staging buffers
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cores = os.cpu_count() or 1
    threads = max(1, min(len(os.sched_getaffinity(0)), cores // max(1, world)))
    ctx = Context(local_rank)
    B = args.batch
    n_windows = N_WINDOWS
    wl = Workload(B * n_windows, seed=args.seed + rank, threads=threads)
    gmap, layers = build_gpu_map(ctx, wl, TARGET_VOXELS, rank, world)
    scans, gts, inits = wl.query_set(layers[-1][0])
    fps = [wl.fp] * B
    owners = [capi.IcpParamsOwner(sigma=SIGMA) for _ in range(B)]
    params = [o.p for o in owners]
    windows = [list(range(w * B, (w + 1) * B)) for w in range(n_windows)]
    resident = [ctx.upload_batch([scans[i] for i in win]) for win in windows]
    # pinned host copies for the e2e leg: x, y, z packed (12 B per point).  The path reads no other channel (the 4th lane of
    # a KITTI .bin is intensity), and packed xyz is also what an MRPT caller holds (CPointsMap keeps x / y / z vectors).
    host = []
    for win in windows:
        flat, offs, stride = Context._concat([np.ascontiguousarray(scans[i][:, :3]) for i in win])
        t = torch.from_numpy(flat).pin_memory()
        host.append((t, offs, stride, t.numpy()))
    ext = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for s in range(warmup):
            fn(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(ext)
        for s in range(steps):
            fn(warmup + s)
        e1.record(ext)
        barrier()
        wall = time.perf_counter() - w0
        ms = max(e0.elapsed_time(e1), 0.0)
        ms = max(ms, 0.0)
        t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1])

    last = {}

    def step_resident(s):
        w = s % n_windows
        last["res"] = ctx.scan_register_batch_resident(gmap, resident[w], fps, inits[windows[w]], params)
        last["w"] = w

    def stage(s):
        t, offs, stride, arr = host[s % n_windows]
        ctx.stage_upload_async(s % 2, arr, offs, stride)

    def step_e2e(s):
        # pipelined C-ABI path: enqueue the H2D of step s+1 (pinned host buffer, copy stream), then run step s
        # on the batch staged earlier; every timed step therefore contains one upload and one compute.
        stage(s + 1)
        w = s % n_windows
        last["res"] = ctx.scan_register_batch_staged(gmap, s % 2, fps, inits[windows[w]], params)
        last["w"] = w

    # ---- timed region 1: resident inputs (value)
    for s in range(args.warmup):
        step_resident(s)
    log(f"[bench] kernel launches before the timed region: {ctx.launch_count}")
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    cuprof = os.environ.get("MLO_BENCH_CUPROF") == "1"   # ncu --profile-from-start off: capture the timed region only
    if cuprof:
        torch.cuda.cudart().cudaProfilerStart()
    ms_dev, ms_wall = timed(step_resident, args.steps, 0)
    if cuprof:
        torch.cuda.cudart().cudaProfilerStop()
    launches = ctx.launch_count - l0
    clocks = sampler.stop()
    # ---- roofline pass: the same K steps again with per-kernel CUDA events inside the library (on its stream).
    # While profiling the library runs the batch as ONE launch sequence, so the match kernel is timed alone on the
    # device; the value pass above overlaps two half-batch sequences on two streams (DESIGN.md "launch policy").
    ctx.profile_enable(True)
    step_resident(0)
    ctx.profile_get(reset=True)
    ms_dev_prof, ms_wall_prof = timed(step_resident, args.steps, 0)
    prof = ctx.profile_get(reset=True)
    ctx.profile_enable(False)
    ms_step = max(ms_dev, ms_wall) / args.steps   # the call returns only after its D2H: wall >= device time
    value = world * B / (ms_step * 1e-3)

    # accuracy of the last batch against ground truth and (below) against the oracle
    res = last["res"]
    win = windows[last["w"]]
    from oracle import oracle_py as O  # checker + CPU baseline leg only
    err_gt = [O.pose_error(r.pose, gts[i]) for r, i in zip(res, win)]
    iters = [int(r.n_iterations) for r in res]

    # ---- timed region 2: e2e through the C ABI with pinned host buffers
    stage(0)
    ms_dev2, ms_wall2 = timed(step_e2e, args.steps, args.warmup)
    ms_step2 = max(ms_dev2, ms_wall2) / args.steps
    e2e_value = world * B / (ms_step2 * 1e-3)
    res_e2e, win_e2e = last["res"], windows[last["w"]]
    h2d = int(np.mean([h[0].numel() * 4 for h in host]))  # bytes of one step's raw clouds
    d2h = int(B * C.sizeof(capi.IcpResult))

    # ---- roofline of the fused NN+residual kernel (SURVEY.md §8(d) byte formula)
    peaks = {}
    try:
        peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    qi, P, nb = prof.nn_query_iterations, prof.nn_candidate_points, prof.nn_blocks
    alg_bytes = 16 * qi + 27 * 16 * qi + 16 * P + 27 * 8 * nb
    nn_ms = prof.nn_kernel_ms
    achieved = (alg_bytes / 1e9) / (nn_ms * 1e-3) if nn_ms > 0 else 0.0
    traffic, traffic_note = None, None
    try:
        tfile = ROOT / "profiles" / "r02_traffic.json"
        tj = json.loads((tfile if tfile.exists() else ROOT / "profiles" / "r01_traffic.json").read_text())
        traffic = tj["dram_bytes_per_launch"]
        traffic_note = (f"ncu --set full capture of one full-activity launch at B=512 ({tj['capture'].split(' ')[0]}); "
                        f"algorithmic bytes of that launch: {tj['algorithmic_bytes_same_launch']:.3g}")
    except Exception:
        pass
    persistent_used = prof.nn_kernel_launches <= 2 * args.steps
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_note": traffic_note,
                "kernel": "k_icp_persistent" if persistent_used else "k_match_accumulate_wl4 (+ k_icp_persistent for the tail)",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6650 GB/s",
                "bytes_per_launch": alg_bytes / max(1, prof.nn_kernel_launches),
                "avg_launch_us": 1e3 * nn_ms / max(1, prof.nn_kernel_launches),
                "launches": int(prof.nn_kernel_launches), "kernel_share_of_step": nn_ms / max(ms_dev_prof, 1e-9),
                "timed_in": "second pass over the same K steps with per-kernel events, single launch sequence",
                "ms_per_step_of_that_pass": max(ms_dev_prof, ms_wall_prof) / args.steps,
                "candidates_per_query": P / max(1, qi)}

    # ---- parity of the timed path (rank 0): EVERY scan of the last timed batch against the oracle, tolerance asserted
    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same scans with the oracle
    cpu = None
    parity = None
    if rank == 0 and not args.no_cpu_baseline:
        omap = build_oracle_map(wl, layers)
        ip_chk = [capi.IcpParamsOwner(sigma=SIGMA) for _ in range(cores)]

        def chk(a):
            j, (r, i) = a
            orr, _ = O.scan_register(omap, scans[i], wl.fp, inits[i], ip_chk[j % cores].p)
            return O.pose_error(r.pose, orr.pose) + (int(r.n_iterations) - int(orr.n_iterations), int(r.termination) - int(orr.termination))
        parity = {"tolerance": "1e-3 m / 1e-2 deg per scan (north_star), asserted"}
        with ThreadPoolExecutor(cores) as ex:
            for leg, (rr, ww) in {"resident": (res, win), "e2e": (res_e2e, win_e2e)}.items():
                deltas = list(ex.map(chk, enumerate(zip(rr, ww))))
                parity[leg] = {"scans": len(deltas), "max_trans_m": max(d[0] for d in deltas),
                               "max_rot_deg": max(d[1] for d in deltas),
                               "iteration_count_differs": int(sum(1 for d in deltas if d[2] != 0)),
                               "termination_differs": int(sum(1 for d in deltas if d[3] != 0))}
                assert parity[leg]["max_trans_m"] <= 1e-3 and parity[leg]["max_rot_deg"] <= 1e-2, \
                    f"GPU result ({leg} leg) differs from the oracle: {parity[leg]}"
        if world == 1:
            v1, n1, dt1, ms3, _ = cpu_scans_per_sec(omap, scans, inits, wl.fp, 1, args.cpu_budget * 0.4, 64, warm_scans=3)
            vN, nN, dtN, _, segN = cpu_scans_per_sec(omap, scans, inits, wl.fp, cores, args.cpu_budget * 0.6, len(scans),
                                                     warm_scans=max(3, 2 * cores))
            cpu = {"value": vN, "unit": "scans/s", "cores": cores, "kind": "port",
                   "sample": f"{nN} scans of this workload in {dtN:.1f}s, one persistent pool of {cores} workers "
                             f"(independent scans in flight, like eval/cli_kitti.sh), warm-up scans untimed",
                   "median_of_5_segments": segN,
                   "single_thread": {"value": v1, "scans": n1,
                                     "ms_filter_1st/run_icp/update_local_map": [float(x) for x in ms3]},
                   "note": "CPU = our restatement of mp2p_icp/mola_metric_maps, not the upstream binary"}

    # ---- sub-records: BASELINE.json configs[2] / [3] / [4] (whole sequences through the full odometry loop)
    subs = {}
    mode = args.sub_records if args.sub_records != "auto" else ("all" if world == 1 else "fleet")
    if mode != "none":
        plan = [("sequences_fleet", "sequence", args.sub_sequences, args.sub_scans)]
        if mode == "all":
            plan += [("sequence_single", "sequence", 1, 2 * args.sub_scans), ("sequence_pt2pl_fleet", "sequence_pt2pl", 8, args.sub_scans // 2),
                     ("ndt_o128_fleet", "ndt", 8, args.sub_scans // 4)]
        for name, kind, S_, N_ in plan:
            run_fleet_workload(ctx, kind, S_, min(N_, 12), rank, world, cores, False, 0)                       # warm-up
            barrier()
            rec = run_fleet_workload(ctx, kind, S_, N_, rank, world, cores, rank == 0 and world == 1 and not args.no_cpu_baseline,
                                     min(N_, args.cpu_scans))
            tt = torch.tensor([rec["wall_s"]], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            rec["value"] = world * S_ * N_ / float(tt[0])      # whole job: all ranks' sequences / slowest rank
            rec["n_gpus"] = world
            if "cpu_baseline" in rec:
                rec["speedup_vs_cpu"] = rec["value"] / rec["cpu_baseline"]["value"]
            subs[name] = rec
            log(f"[bench] {name}: {rec['value']:.0f} scans/s" + (f", CPU {rec['cpu_baseline']['value']:.0f} scans/s" if "cpu_baseline" in rec else ""))

    if rank == 0:
        line = {"metric": "scans/sec", "value": value, "unit": "scans/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32 distances / f64 normal equations", "data": "synthetic",
                "config": config_dict(B, scans, gmap.stats(), world),
                "e2e": {"value": e2e_value, "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_step2, "host_layout": "x,y,z float32 packed (12 B/pt), pinned, NUMA-local to the GPU",
                        "host_link_GBps": h2d / (ms_step2 * 1e-3) / 1e9, "numa": numa},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "quality": {"mean_iterations": float(np.mean(iters)), "max_err_vs_gt_m": max(e[0] for e in err_gt),
                            "median_err_vs_gt_m": float(np.median([e[0] for e in err_gt])), "parity_vs_oracle": parity},
                "timing": {"device_ms_total": ms_dev, "wall_ms_total": ms_wall},
                "sub_records": subs}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()
