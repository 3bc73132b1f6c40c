"""Seeded synthetic Lyft-shaped inputs for tests and benchmarks.

No dataset is reachable offline, so sweeps, ground-truth boxes, the per-slot data mean and the
PFN weights are generated.  Everything is a pure function of its seed.
"""
import numpy as np

from .config import PPConfig, class_dims

# rough Lyft class frequencies (car-heavy), order = PPConfig.class_names
_CLASS_P = np.array([0.002, 0.03, 0.015, 0.80, 0.001, 0.003, 0.05, 0.04, 0.059])
_CLASS_P = _CLASS_P / _CLASS_P.sum()


def make_sweep(seed, n_sweeps=1, ego_shift=0.3):
    """One lidar cloud as float32 [N,5] = (x, y, z, intensity, ring), like a Lyft .bin file.

    64-ring spinning lidar 1.8 m above a ground plane: elevations linspace(-25deg, +3deg, 64),
    1250 azimuth steps with N(0,1e-3) jitter, range min(1.8/sin(-el), 100 m), ~40 occluder
    sectors at 4-50 m, 0.2 % range noise, 5 % dropout => ~67 k points.  ``n_sweeps`` > 1
    concatenates sweeps (seed, seed+1, ...) shifted by ``ego_shift`` metres each, like the
    reference's multi-sweep aggregation (data/dataset.py:58-85)."""
    clouds = []
    for s in range(n_sweeps):
        rng = np.random.default_rng(seed + s)
        el = np.deg2rad(np.linspace(-25.0, 3.0, 64))
        az = np.linspace(-np.pi, np.pi, 1250, endpoint=False)
        el_g, az_g = np.meshgrid(el, az, indexing="ij")
        az_g = az_g + rng.normal(0.0, 1e-3, az_g.shape)
        with np.errstate(divide="ignore"):
            rng_ground = np.where(el_g < 0, 1.8 / np.sin(-el_g), 100.0)
        r = np.minimum(rng_ground, 100.0)
        # occluders: angular sectors that return early (vehicles, walls)
        n_occ = 40
        occ_az = rng.uniform(-np.pi, np.pi, n_occ)
        occ_w = rng.uniform(0.01, 0.12, n_occ)
        occ_r = rng.uniform(4.0, 50.0, n_occ)
        occ_h = rng.uniform(0.5, 3.5, n_occ)
        for a0, w, rr, hh in zip(occ_az, occ_w, occ_r, occ_h):
            d = np.abs(np.angle(np.exp(1j * (az_g - a0))))
            z_at = 1.8 + rr * np.tan(el_g)
            hit = (d < w) & (rr < r) & (z_at < hh) & (z_at > 0)
            r = np.where(hit, rr, r)
        r = r * (1.0 + rng.normal(0.0, 0.002, r.shape))
        keep = (rng.uniform(size=r.shape) > 0.05) & (r < 99.5)
        x = r * np.cos(el_g) * np.cos(az_g) + s * ego_shift
        y = r * np.cos(el_g) * np.sin(az_g)
        z = 1.8 + r * np.sin(el_g) - 1.8   # sensor frame: ground near z = -1.8 + 1.8 = 0 offset
        z = z - 1.0
        ring = np.broadcast_to(np.arange(64)[:, None], r.shape)
        pts = np.stack([x[keep], y[keep], z[keep], np.full(keep.sum(), 100.0), ring[keep]], axis=1)
        # a spinning lidar emits column by column (azimuth-major), not ring by ring
        order = np.argsort(np.broadcast_to(np.arange(1250)[None, :], r.shape)[keep], kind="stable")
        clouds.append(pts[order])
    return np.concatenate(clouds, axis=0).astype(np.float32)


def make_gt(seed, G, cfg=None):
    """G ground-truth boxes in canvas units.  Returns dict of float64 arrays:
    centers [G,3] (canvas space, NOT y-flipped), wlh [G,3], yaw [G] (radians), cls int32 [G].
    Half of the yaws are snapped near 0 / +-pi/2 and box centres near the anchor lattice so that
    positives (IoU > 0.6) actually occur."""
    cfg = cfg or PPConfig()
    rng = np.random.default_rng(10_000 + seed)
    dims = class_dims(cfg.x_step)
    cls = rng.choice(len(cfg.class_names), size=G, p=_CLASS_P).astype(np.int32)
    wlh = np.stack([dims[cfg.class_names[c]] for c in cls] + [np.zeros(3)])[:G] * rng.uniform(0.8, 1.2, (G, 3))
    centers = np.stack([rng.uniform(20, cfg.canvas_width - 20, G),
                        rng.uniform(20, cfg.canvas_height - 20, G),
                        rng.uniform(-1.0, 2.0, G)], axis=1)
    yaw = rng.uniform(-np.pi, np.pi, G)
    snap = rng.uniform(size=G) < 0.5
    snapped = rng.choice([0.0, np.pi / 2, -np.pi / 2, np.pi - 1e-3], size=G) + rng.normal(0, 0.03, G)
    yaw = np.where(snap, snapped, yaw)
    return {"centers": centers, "wlh": wlh, "yaw": yaw, "cls": cls}


def make_data_mean(P, N, seed=0, dense=True):
    """Synthetic stand-in for pillar_means.pkl (make_means.py:28-37): float32 [9*P*N], one mean
    per (d,p,n) slot.  ``dense`` gives every slot a non-zero mean (worst case for the PFN);
    otherwise slots beyond a plausible occupancy are exactly zero like never-filled slots are."""
    rng = np.random.default_rng(20_000 + seed)
    occ = np.exp(-np.arange(N) / 6.0)[None, :] * np.exp(-np.arange(P) / (P / 3.0))[:, None]
    base = np.array([0.5, -0.3, -0.9, 80.0, 250.0, 260.0, 0.01, -0.01, 0.02])
    m = base[:, None, None] * occ[None] + rng.normal(0, 0.05, (9, P, N)) * occ[None]
    if dense:
        m = m + rng.normal(0, 1e-3, (9, P, N))
    else:
        m = np.where(occ[None] > 1e-3, m, 0.0)
    return m.astype(np.float32).reshape(-1)


def make_pfn_params(seed=0, in_channels=9, out_channels=64, flip_gamma=False):
    """PFN parameters with nn.Conv2d / nn.BatchNorm2d default-style init, as numpy float32."""
    rng = np.random.default_rng(30_000 + seed)
    bound = 1.0 / np.sqrt(in_channels)
    p = {
        "conv_w": rng.uniform(-bound, bound, (out_channels, in_channels)).astype(np.float32),
        "conv_b": rng.uniform(-bound, bound, out_channels).astype(np.float32),
        "bn_w": np.ones(out_channels, np.float32),
        "bn_b": np.zeros(out_channels, np.float32),
        "running_mean": np.zeros(out_channels, np.float32),
        "running_var": np.ones(out_channels, np.float32),
    }
    if flip_gamma:
        p["bn_w"] = (rng.uniform(0.5, 1.5, out_channels) * rng.choice([-1.0, 1.0], out_channels)).astype(np.float32)
        p["bn_b"] = rng.normal(0, 0.1, out_channels).astype(np.float32)
        p["running_mean"] = rng.normal(1.0, 0.5, out_channels).astype(np.float32)
        p["running_var"] = rng.uniform(0.5, 4.0, out_channels).astype(np.float32)
    return p
