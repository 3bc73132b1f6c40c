#!/usr/bin/env python
"""Benchmark of the B200 PointPillars input path (BASELINE.json metric: sweeps/sec for
pillarize + PFN + IoU targets, with the HBM roofline of the dominant kernel).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference CPU path (rank 0)

One step = one pass of the hot path over one batch of 4 synthetic Lyft-shaped sweeps per GPU
(BASELINE.json configs[1] + configs[2]: batch-4 pillarize + PFN/scatter, 100 GT boxes per sweep
against the full 540000-anchor grid).  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import socket
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

BATCH = 4
N_GT = 100
METRIC = "sweeps/sec (pillarize+PFN+IoU targets)"
UNIT = "sweeps/s"
WORKLOAD = ("batch-%d synthetic Lyft-shaped sweeps (~67k pts x5 f32), config.py grid 600x600, P=24000 N=200 "
            "D=9 C=64: pillarize+decorate+data_mean -> PFN (train-mode BN) + scatter to [64,600,600] -> "
            "IoU/target encode, %d GT boxes vs 540000 anchors" % (BATCH, N_GT))


# ------------------------------------------------------------------------------------------------
def reduce_over_ranks(ms_local, units_local, device):
    """MAX of the per-rank time, SUM of the per-rank units (the only collectives of this path)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(ms_local), float(units_local)
    t = torch.tensor([ms_local], dtype=torch.float64, device=device)
    u = torch.tensor([units_local], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t.item()), float(u.item())


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML while the timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:  # noqa: BLE001
            return local
    return local


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU path (oracle): used for cpu_baseline and for --impl reference.  Never on the product path.
def _cpu_state():
    import numpy as np
    import torch
    from oracle import targets as T
    from pp_b200 import synth
    st = getattr(_cpu_state, "cache", None)
    if st is None:
        corners, centers, wlh, yaw = T.anchor_arrays()
        prm = synth.make_pfn_params(0)
        st = {
            "anchors": (corners, centers, T.LazyAnchorBoxes(centers, wlh, yaw)),
            "mean": torch.from_numpy(synth.make_data_mean(24000, 200, seed=0, dense=True)),
            "prm": {k: torch.from_numpy(v.copy()) for k, v in prm.items()},
        }
        _cpu_state.cache = st
    return st


def cpu_path_one_sweep(seed):
    """The reference's per-sample path on one host core, restated (oracle/): create_pillars +
    dataset glue (data/dataset.py:88-106), PPFeatureNet + PPScatter with the float32 library ops
    of model/model.py:31-62, boxes_to_image_space + create_target (utils/box_utils.py).  With
    oracle/_ref present, create_pillars and make_ious are the reference's own compiled code."""
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import config as ocfg, glue, pfn, ref, targets as T
    from pp_b200 import synth
    st = _cpu_state()
    m = ref.load()
    s = synth.make_sweep(seed)
    gt = synth.make_gt(seed, N_GT)
    t0 = time.perf_counter()
    lidar = np.ascontiguousarray(s[:, :4].astype(np.float64).T).T        # [4,N].T view like dataset.py:88
    x, inds = glue.pillarize(lidar, st["mean"], create_pillars=(m.create_pillars if m else None))
    p = st["prm"]
    rm, rv = p["running_mean"].clone(), p["running_var"].clone()
    with torch.no_grad():
        y = pfn.reference_forward_f32(x[None], p["conv_w"], p["conv_b"], p["bn_w"], p["bn_b"], rm, rv, True)
        canvas = pfn.scatter(y, inds[None], 600, 600)
    boxes = [T.Box(gt["centers"][i], gt["wlh"][i], gt["yaw"][i], ocfg.CLASS_NAMES[int(gt["cls"][i])])
             for i in range(len(gt["yaw"]))]
    gc, gcor = T.boxes_to_image_space(boxes)
    corners, centers, lazy = st["anchors"]
    cls, reg = T.create_target(corners, gcor, centers, gc, lazy, boxes,
                               make_ious=(m.make_ious if m else None))
    c_t = torch.from_numpy(cls).float(); r_t = torch.from_numpy(reg).float()
    dt = time.perf_counter() - t0
    return dt, float(canvas.sum()) + float(c_t.sum()) + float(r_t.sum())


def cpu_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        n = os.cpu_count() or 1
    try:
        import psutil
        by_mem = int(psutil.virtual_memory().available / (7 * 2 ** 30))
        n = max(1, min(n, by_mem))
    except Exception:  # noqa: BLE001
        pass
    return n


def cpu_kind():
    from oracle import ref
    return "reference" if ref.load() is not None else "port"


def cpu_kind_note():
    if cpu_kind() == "reference":
        return ("create_pillars/make_ious = the reference's data/pillars.cpp compiled against the Boost stand-in "
                "(oracle/boost_shim; Boost absent from the image); glue/PFN/create_target = restated Python "
                "(numpy / torch CPU float32)")
    return "oracle port (oracle/pp_oracle.c + numpy / torch CPU float32)"


def run_cpu_pool(seeds, cores):
    """Process the given sweeps on `cores` worker processes; returns wall seconds."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    if cores <= 1:
        for s in seeds:
            cpu_path_one_sweep(s)
    else:
        with ctx.Pool(min(cores, len(seeds))) as pool:
            pool.map(cpu_path_one_sweep, seeds, chunksize=1)
    return time.perf_counter() - t0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import native
    native.build()
    cores = cpu_cores()
    # calibrate: one sweep on one core
    dt1, _ = cpu_path_one_sweep(0)
    est_parallel = dt1 * 1.6                      # memory-bound stages slow down when all cores run
    sample = max(1, cores)                        # sweeps per step: one per worker
    budget_s = 240.0
    steps, warmup = args.steps, args.warmup
    while (steps + warmup) * est_parallel > budget_s and steps > 1:
        steps = max(1, steps // 2)
    warmup = min(warmup, 1) if (steps + warmup) * est_parallel > budget_s else warmup
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    seeds = iter(range(1000, 100000))
    with ctx.Pool(cores) as pool:
        for _ in range(warmup):
            pool.map(cpu_path_one_sweep, [next(seeds) for _ in range(sample)], chunksize=1)
        t0 = time.perf_counter()
        for _ in range(steps):
            pool.map(cpu_path_one_sweep, [next(seeds) for _ in range(sample)], chunksize=1)
        wall = time.perf_counter() - t0
    value = sample * steps / wall
    sample_txt = ("%d steps x %d sweeps (one per worker process), each sweep = full per-sample CPU path of the "
                  "workload; %s; single-sweep single-core latency %.2f s" % (steps, sample, cpu_kind_note(), dt1))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * wall / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "host": socket.gethostname()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": cpu_kind(), "sample": sample_txt},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def algorithmic_bytes(B, P, N, C, H, W, A, K, total_points, has_mean):
    """Algorithmic (compulsory) HBM bytes per launch of each streaming kernel."""
    x_bytes = 9 * P * N * 4
    return {
        "k_emit_dense": B * x_bytes + (x_bytes if has_mean else 0) + total_points * 16 + B * P * 24,
        "k_pfn_stats": B * x_bytes + B * P * 2 * C * 4,
        "k_pfn_stats_tc": B * x_bytes + B * P * 2 * C * 4,
        "k_canvas": B * C * H * W * 4 + B * H * W * 4,
        # fused path: data_mean read once + the ext rows of the live pillars (2 padding fields per sweep)
        "k_pfn_pad_tc": (x_bytes if has_mean else 0) + B * 16500 * 2 * C * 4,
        "k_encode_zero": 2 * B * A * 9 * 4,  # zero stream of cls [B,A,9] and reg [B,A,9] (flagged rows patched after)
        "k_encode": B * A * 9 * 4,
    }


def training_rows(x, inds, targets, dev, peak, steps, path=None, d_pts=None, offsets=None, gt=None, fused_path=None):
    """PFN + scatter backward (pp_pfn_backward) and the loss front-end (pp_loss) on the step's own x / inds /
    targets, network outputs random: per-kernel CUDA-event times and, for the streaming loss kernels, GB/s."""
    import torch
    from pp_b200 import _lib, loss as pl, model as pm
    L = _lib.load()
    B = x.shape[0]
    cls_t, reg_t = targets
    net = pm.PPFeatureScatter(9, 64).to(dev).train()
    lossm = pl.PPLoss(0, 1, 250, 2, dev)
    g_canvas = torch.randn((B, 64, net.canvas_height, net.canvas_width), device=dev)
    cls = torch.randn((B, 54, 300, 300), device=dev) * 1.5 - 3.0
    reg = torch.randn((B, 48, 300, 300), device=dev)

    pts_copy = d_pts.clone() if d_pts is not None else None
    rigid = None
    if pts_copy is not None:
        import numpy as np
        rigid = np.stack([[[np.cos(0.01 * k), -np.sin(0.01 * k), 0, 0.4 * k], [np.sin(0.01 * k), np.cos(0.01 * k), 0, 0.0],
                           [0, 0, 1, 0.0]] for k in range(len(offsets) - 1)])
        rigid = torch.from_numpy(rigid.reshape(-1, 12)).to(dev)
        offs_dev = torch.tensor(offsets, dtype=torch.int64, device=dev)

    def one():
        net.zero_grad(set_to_none=True)
        net(x, inds).backward(g_canvas)
        c = cls.clone().requires_grad_(True)
        r = reg.clone().requires_grad_(True)
        lossm(c, r * 1.0, cls_t, reg_t)[4].backward()
        if pts_copy is not None:
            pts_copy.copy_(d_pts)
            path.aggregate(pts_copy, offs_dev, rigid)
        if fused_path is not None:                   # the x-free path made trainable: forward + sparse backward
            fused_path.net.zero_grad(set_to_none=True)
            fused_path.pillarize_encode_train(d_pts, offsets)[0].backward(g_canvas)
        if gt is not None:                           # the same targets as a positives list, and the loss fed by it
            path.targets_as_list = True
            pos = path.targets(gt[0], gt[1])[0]
            path.targets_as_list = False
            c = cls.clone().requires_grad_(True)
            r = reg.clone().requires_grad_(True)
            lossm(c, r * 1.0, pos)[4].backward()

    one()
    torch.cuda.synchronize()
    L.pp_profile_enable(1)
    for _ in range(steps):
        one()
    rep = _lib.profile_report()
    L.pp_profile_enable(0)
    n_cls, n_reg = cls.numel() * 4, reg.numel() * 4
    alg = {"k_loss_cls_tma": 4 * n_cls + 2 * B * 90000 * 4, "k_loss_cls_list": 3 * n_cls + 2 * B * 90000 * 4,
           "k_loss_reg": reg_t.numel() * 4}
    if d_pts is not None:
        alg["k_aggregate"] = d_pts.shape[0] * (3 * 4 * 2 + 8)          # x,y,z read + written; sector granularity not counted
    rows = {}
    for name, (n, ms) in rep.items():
        if not (name.startswith("k_loss") or name.startswith("k_pfn_bwd") or name.startswith("k_pos_") or
                name == "k_aggregate"):
            continue
        k = {"us_per_launch": ms * 1e3 / n, "launches_per_step": n / steps}
        if name in alg:
            k["alg_bytes_per_launch"] = alg[name]
            k["GBps"] = alg[name] / (ms / n * 1e-3) / 1e9
            k["frac_of_peak"] = k["GBps"] / peak
        rows[name] = k
    slots = float(x.shape[0] * x.shape[2] * x.shape[3])
    if "k_pfn_bwd" in rows:
        rows["k_pfn_bwd"]["note"] = ("FP32-issue bound, not HBM: 30 FMA per (slot, channel) for z and the BatchNorm moment "
                                     "matrices; %.1f G(slot*channel)/s" % (slots * 64 / rows["k_pfn_bwd"]["us_per_launch"] / 1e3))
    if "k_pfn_bwd_pad" in rows:
        rows["k_pfn_bwd_pad"]["note"] = ("pp_input_path_backward, pass A: padding-slot moments + per-sweep suffix arg-max once over "
                                         "[P,N]; k_pfn_bwd_live: the live pillars from the compact state (x never built)")
    return {"what": "pp_input_path_backward through InputPath.pillarize_encode_train (k_pfn_bwd_pad/_live: sparse formulation); "
                    "pp_pfn_backward through PPFeatureScatter.backward (training-mode BatchNorm) and pp_loss through "
                    "PPLoss forward + backward from the dense targets and from the positives list of pp_assign_targets_list "
                    "(k_pos_*, k_loss_*_list), batch of %d sweeps; pp_aggregate_sweeps (rigid transform + remove_close, in "
                    "place) on the batch's raw points" % B, "kernels": rows}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this path has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pp_b200
    from pp_b200 import _lib, pipeline, synth
    L = _lib.load()
    if getattr(args, "reserve_sms", None) is not None:
        L.pp_set_option(b"pad_reserve_sms", int(args.reserve_sms))
    cfg = pp_b200.PPConfig()
    P, N, C, H, W = cfg.max_pillars, cfg.max_points_per_pillar, cfg.feature_net_out, cfg.canvas_height, cfg.canvas_width
    mean = synth.make_data_mean(P, N, seed=0, dense=True)
    fused = not args.dense_path
    path = pipeline.InputPath(cfg, device=dev, data_mean=mean, pfn_params=synth.make_pfn_params(0), training=True,
                              fused=fused)
    anchors = path.ensure_anchors()
    A = anchors.A
    # each rank owns its own batch of BATCH sweeps per step (weak scaling, no data-path collective)
    sweeps = [synth.make_sweep(rank * BATCH + i) for i in range(BATCH)]
    gts = [synth.make_gt(rank * BATCH + i, N_GT) for i in range(BATCH)]
    batch = path.pack_host_batch(sweeps, gts)
    T = batch["offsets"][-1]
    d_pts, gt_dev = path.upload(batch)
    out = {
        "pillars": (torch.empty((BATCH, 9, P, N), dtype=torch.float32, device=dev) if not fused else None,
                    torch.empty((BATCH, P, 3), dtype=torch.int64, device=dev),
                    torch.empty(BATCH, dtype=torch.int32, device=dev)),
        "canvas": torch.empty((BATCH, C, H, W), dtype=torch.float32, device=dev),
        "targets": (torch.empty((BATCH, A, cfg.num_classes), dtype=torch.float32, device=dev),
                    torch.empty((BATCH, A, 9), dtype=torch.float32, device=dev)),
    }

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1)

    # streaming loop: consecutive steps alternate between the two lanes of the InputPath (and between two
    # sets of output buffers), at most two steps in flight; every step is a full pass over its batch
    outs = [out, {k: (tuple(torch.empty_like(t) if t is not None else None for t in v) if isinstance(v, tuple)
                      else torch.empty_like(v)) for k, v in out.items()}]
    inflight = []
    tick = [0]

    def step_dev():
        h = path.step_device_async(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out=outs[tick[0] & 1])
        tick[0] += 1
        inflight.append(h)
        if len(inflight) > args.inflight - 1:
            inflight.pop(0).synchronize()

    def drain_dev():
        while inflight:
            inflight.pop(0).synchronize()

    pending = []

    def step_e2e():
        # same loop from pinned HOST buffers: the H2D copy and the pillarize stage of this step overlap the
        # encode stage of the previous one; every step's counters are read back on the host inside the timed
        # region
        pending.append(path.step_host_async(batch, out=outs[tick[0] & 1]))
        tick[0] += 1
        if len(pending) > args.inflight - 1:
            pending.pop(0).counters()

    def drain_e2e():
        while pending:
            pending.pop(0).counters()

    for _ in range(max(args.warmup, 3)):
        step_dev()
    drain_dev()
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    n0 = L.pp_launch_count()

    def timed_loop():
        for _ in range(args.steps):
            step_dev()
        drain_dev()

    ms = timed(timed_loop, 1)
    launches = (L.pp_launch_count() - n0)
    clocks = sampler.stop()
    ms_max, units = reduce_over_ranks(ms, float(BATCH * args.steps), dev)
    value = units / (ms_max / 1e3)

    # the same device loop with K3 emitting the positives list instead of the dense [B,A,9] tensors
    list_mode = None
    if fused and not args.no_dense_reference:
        path.targets_as_list = True
        for _ in range(3):
            step_dev()
        drain_dev()
        ms_l = timed(timed_loop, 1)
        path.targets_as_list = False
        ms_l_max, units_l = reduce_over_ranks(ms_l, float(BATCH * args.steps), dev)
        list_mode = {"what": "same streaming loop, targets as a positives list (pp_assign_targets_list): K3 writes a few KB "
                             "instead of 155 MB per step; consumer pp_loss_list", "value": units_l / (ms_l_max / 1e3),
                     "unit": UNIT, "ms_per_step": ms_l_max / args.steps}

    # end-to-end through the host-facing call: pinned host buffers in, counters out, every step
    for _ in range(3):
        step_e2e()
    drain_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_host0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step_e2e()
    drain_e2e()                                # the last step's counters are read before the clock stops
    e1.record()
    barrier()
    ms_e = max(e0.elapsed_time(e1), 0.0)
    ms_e_host = (time.perf_counter() - t_host0) * 1e3
    ms_e_max, units_e = reduce_over_ranks(ms_e, float(BATCH * args.steps), dev)
    e2e_value = units_e / (ms_e_max / 1e3)
    h2d = int(batch["blob"].numel())
    d2h = int(BATCH * 4 + BATCH * 4 * 4)

    alg = algorithmic_bytes(BATCH, P, N, C, H, W, A, cfg.num_classes, T, True)
    peak, peak_src = measured_peak()

    def kernel_table(pth, o):
        """Per-kernel durations, live, CUDA events on the launching stream (separate instrumented pass,
        stream overlap off so that each kernel is timed running alone)."""
        pth.overlap_targets = False
        L.pp_profile_enable(1)
        barrier()
        for _ in range(args.steps):
            pth.step_device(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out=o)
        rep = _lib.profile_report()
        L.pp_profile_enable(0)
        pth.overlap_targets = True
        ks = []
        tot_ms = sum(v[1] for v in rep.values()) or 1.0
        for name, (n, total_ms) in rep.items():
            k = {"name": name, "launches_per_step": n / args.steps, "ms_per_launch": total_ms / n,
                 "share_of_kernel_time": total_ms / tot_ms}
            if name in alg:
                k["alg_bytes_per_launch"] = alg[name]
                k["GBps"] = alg[name] / (total_ms / n * 1e-3) / 1e9
                k["frac_of_peak"] = k["GBps"] / peak
            ks.append(k)
        ks.sort(key=lambda k: -k["share_of_kernel_time"])
        dom = next((k for k in ks if "GBps" in k), None)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(dom["name"])
        except Exception:  # noqa: BLE001
            pass
        roof = None
        if dom is not None:
            roof = {"bound": "hbm", "kernel": dom["name"], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                    "frac": dom["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                    "alg_bytes_per_launch": dom["alg_bytes_per_launch"], "ms_per_launch": dom["ms_per_launch"]}
        return ks, roof

    kernels, roofline = kernel_table(path, out)

    # the other formulation of the same step, for reference (same inputs, same outputs)
    other = None
    training = None
    if fused and not args.no_dense_reference:
        path2 = pipeline.InputPath(cfg, device=dev, data_mean=mean, pfn_params=synth.make_pfn_params(0),
                                   training=True, fused=False, anchors=anchors)
        out2 = dict(out)
        out2["pillars"] = (torch.empty((BATCH, 9, P, N), dtype=torch.float32, device=dev), out["pillars"][1],
                           out["pillars"][2])
        outs2 = [out2, dict(outs[1])]
        outs2[1]["pillars"] = (torch.empty((BATCH, 9, P, N), dtype=torch.float32, device=dev), outs[1]["pillars"][1],
                               outs[1]["pillars"][2])
        infl2 = []

        def loop2():
            for i in range(args.steps):
                h = path2.step_device_async(d_pts, batch["offsets"], gt_dev, batch["gt_offsets"], out=outs2[i & 1])
                infl2.append(h)
                if len(infl2) > 1:
                    infl2.pop(0).synchronize()
            while infl2:
                infl2.pop(0).synchronize()

        loop2()
        ms2 = timed(loop2, 1)
        ms2_max, units2 = reduce_over_ranks(ms2, float(BATCH * args.steps), dev)
        k2, roof2 = kernel_table(path2, out2)
        other = {"what": "signature-preserving sequence pp_pillarize -> x [B,9,P,N] -> pp_pfn_scatter (x materialised)",
                 "value": units2 / (ms2_max / 1e3), "unit": UNIT, "ms_per_step": ms2_max / args.steps,
                 "roofline": roof2, "kernels": k2}
        # the two training-side rows next to the path, timed on this step's own outputs
        if rank == 0 and world == 1 and not args.no_training_rows:
            training = training_rows(out2["pillars"][0], out2["pillars"][1], out["targets"], dev, peak, args.steps,
                                     path=path2, d_pts=d_pts, offsets=batch["offsets"],
                                     gt=(gt_dev, batch["gt_offsets"]), fused_path=path)
        del out2, path2

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import native
        native.build()
        cores = cpu_cores()
        n_s = max(cores, 2) if cores > 1 else 2
        wall = run_cpu_pool(list(range(500, 500 + n_s)), cores)
        cpu_baseline = {"value": n_s / wall, "unit": UNIT, "cores": cores, "kind": cpu_kind(),
                        "sample": "%d sweeps of the workload on %d worker processes, %.1f s wall; %s" % (
                            n_s, cores, wall, cpu_kind_note())}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (PFN, features) + f64 (binning, means, IoU)",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "sweeps_per_gpu_per_step": BATCH, "points_per_step_per_gpu": int(T),
                       "path": ("fused pp_input_path: pillarize stages -> PFN + scatter straight from the compact "
                                "per-point state, x [B,9,P,N] never materialised (padding slots evaluated once per "
                                "(p,n)); see dense_path for the signature-preserving sequence") if fused else
                               "pp_pillarize -> x [B,9,P,N] -> pp_pfn_scatter",
                       "data_mean": "dense synthetic per-slot mean [9*P*N]", "bn": "training mode",
                       "l2": "no flush: per-step working set ~1.5 GB (x 691 MB, canvas 369 MB, targets 156 MB, "
                             "data_mean 173 MB) >> 126 MB L2",
                       "loop": "streaming: steps alternate between two stream lanes (pillarize stage of step k+1 "
                               "overlaps the encode stage of step k; encode stages ordered), <= 2 steps in flight, "
                               "timed region ends after the last step has completed",
                       "parallelism": "dp%d (one process per GPU, sweeps sharded, no collective)" % world},
            "roofline": roofline, "kernels": kernels, "dense_path": other, "list_targets": list_mode, "training_rows": training,
            "cpu_baseline": cpu_baseline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e_max / args.steps, "host_wall_ms_per_step": ms_e_host / args.steps,
                    "pipeline": "step_host_async: copy stream + 2 staging buffers + 2 kernel lanes, at most two steps "
                                "in flight; each step's counters are read back from pinned memory before the "
                                "next-but-one step is issued"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--reserve-sms", type=int, default=None, help="SMs the padding pass leaves free (experiment)")
    ap.add_argument("--no-training-rows", action="store_true", help="skip the PFN backward / loss front-end timings")
    ap.add_argument("--no-dense-reference", action="store_true", help="skip the extra dense_path measurement")
    ap.add_argument("--inflight", type=int, default=2, help="steps in flight in the streaming loops (>= 2)")
    ap.add_argument("--dense-path", action="store_true",
                    help="time the signature-preserving sequence pp_pillarize -> x [B,9,P,N] -> pp_pfn_scatter "
                         "instead of the fused pp_input_path (x never materialised)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port)] + sys.argv
        os.execv(sys.executable, cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
