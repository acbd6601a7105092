"""Deterministic synthetic LiDAR scans in the KITTI .bin layout (N x 4 float32: x, y, z, intensity).

There is no KITTI / MulRan data in this environment, so every parity test, the bench and the CPU
baseline run on scans drawn from this generator.  A *scene* is a static world of boxes (buildings) and poles (trunks);
a *scan* observes a scene from a sensor pose, so two scans of the same scene with different poses are true loop-closure
candidates that exercise the whole candidate cascade (kNN -> constellation -> pairwise -> GMM-L2).

Generator spec: 14-30 boxes (centre U(-78,78)^2, >= 8 m from the origin, sides U(6,25) m, yaw U(0,pi),
height U(2.5,9) m) + 50-130 poles (radius U(0.3,1.5) m, height U(2,7) m); ground at z = -1.73 m; 45 % of the points on
the ground at range 3 + 57 u^1.5, 55 % on structures, only on box faces whose outward normal faces the sensor; N(0, 0.02)
noise; points ordered by azimuth like a spinning LiDAR sweep.  torch is used as plumbing only (it runs the same code
on the CPU for tests and on the GPU for the bench, where generating 5 000 x 120 000 points on the host would dominate).
"""
import math

import torch

GROUND_Z = -1.73
MAX_BOX = 30
MAX_POLE = 130
N_EMIT = MAX_BOX * 4 + MAX_POLE


def _scene_params(seed: int) -> torch.Tensor:
    """[N_EMIT, 8] rows: (kind, ax, ay, bx, by, height, nx, ny); kind 0 = unused, 1 = box face (a -> b, outward normal
    n), 2 = pole (centre a, radius bx)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(0x5CE7E000 + int(seed))
    out = torch.zeros(N_EMIT, 8, dtype=torch.float64)
    n_box = int(torch.randint(14, MAX_BOX + 1, (1,), generator=g))
    n_pole = int(torch.randint(50, MAX_POLE + 1, (1,), generator=g))
    k = 0
    for _ in range(n_box):
        while True:
            c = (torch.rand(2, generator=g, dtype=torch.float64) * 2 - 1) * 78
            if float(c.norm()) >= 8.0:
                break
        sides = 6 + 19 * torch.rand(2, generator=g, dtype=torch.float64)
        yaw = float(torch.rand(1, generator=g, dtype=torch.float64)) * math.pi
        h = 2.5 + 6.5 * float(torch.rand(1, generator=g, dtype=torch.float64))
        ux, uy = math.cos(yaw), math.sin(yaw)
        hx, hy = float(sides[0]) / 2, float(sides[1]) / 2
        corners = []
        for sx, sy in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
            corners.append((float(c[0]) + sx * hx * ux - sy * hy * uy, float(c[1]) + sx * hx * uy + sy * hy * ux))
        for f in range(4):
            a, b = corners[f], corners[(f + 1) % 4]
            mx, my = (a[0] + b[0]) / 2 - float(c[0]), (a[1] + b[1]) / 2 - float(c[1])
            nn = math.hypot(mx, my)
            out[k] = torch.tensor([1, a[0], a[1], b[0], b[1], h, mx / nn, my / nn], dtype=torch.float64)
            k += 1
    k = MAX_BOX * 4
    for _ in range(n_pole):
        c = (torch.rand(2, generator=g, dtype=torch.float64) * 2 - 1) * 78
        r = 0.3 + 1.2 * float(torch.rand(1, generator=g, dtype=torch.float64))
        h = 2.0 + 5.0 * float(torch.rand(1, generator=g, dtype=torch.float64))
        out[k] = torch.tensor([2, float(c[0]), float(c[1]), r, 0, h, 0, 0], dtype=torch.float64)
        k += 1
    return out


_SCENE_CACHE = {}


def scene_params(seed: int) -> torch.Tensor:
    if seed not in _SCENE_CACHE:
        _SCENE_CACHE[seed] = _scene_params(seed)
    return _SCENE_CACHE[seed]


def sensor_pose(scene_seed: int, visit: int):
    """Visit 0 sits at the scene origin with yaw 0; revisits are perturbed by U(-3,3) m and U(-pi,pi)."""
    if visit == 0:
        return 0.0, 0.0, 0.0
    g = torch.Generator(device="cpu")
    g.manual_seed(0x9051E000 + 7919 * int(scene_seed) + int(visit))
    u = torch.rand(3, generator=g, dtype=torch.float64)
    return float(u[0] * 6 - 3), float(u[1] * 6 - 3), float(u[2] * 2 * math.pi - math.pi)


def make_scans(scene_seeds, visits, n_pts: int = 120000, device="cpu", noise_seed: int = 0) -> torch.Tensor:
    """Returns float32 [B, n_pts, 4] on `device`; scan b observes scene `scene_seeds[b]` at visit `visits[b]`."""
    B = len(scene_seeds)
    dev = torch.device(device)
    emit = torch.stack([scene_params(int(s)) for s in scene_seeds]).to(dev)  # [B, E, 8] f64
    poses = torch.tensor([sensor_pose(int(s), int(v)) for s, v in zip(scene_seeds, visits)], dtype=torch.float64,
                         device=dev)  # [B, 3]
    g = torch.Generator(device=dev)
    g.manual_seed(0x2024_0925 + 1000003 * int(noise_seed) + 31 * int(scene_seeds[0]) + int(visits[0]))
    n_ground = int(0.45 * n_pts)
    n_struct = n_pts - n_ground

    sx, sy, th = poses[:, 0:1], poses[:, 1:2], poses[:, 2:3]
    kind = emit[:, :, 0]
    ax, ay, bx, by, hh, nx, ny = (emit[:, :, i] for i in range(1, 8))
    # weights: faces = length * height / range (only faces whose outward normal faces the sensor); poles = 6 r h / range
    fcx, fcy = (ax + bx) / 2, (ay + by) / 2
    flen = torch.sqrt((bx - ax) ** 2 + (by - ay) ** 2)
    frange = torch.sqrt((fcx - sx) ** 2 + (fcy - sy) ** 2).clamp_min(1.0)
    facing = ((sx - fcx) * nx + (sy - fcy) * ny) > 0
    w_face = torch.where((kind == 1) & facing, flen * hh / frange, torch.zeros_like(flen))
    prange = torch.sqrt((ax - sx) ** 2 + (ay - sy) ** 2).clamp_min(1.0)
    w_pole = torch.where(kind == 2, 6 * bx * hh / prange, torch.zeros_like(flen))
    w = (w_face + w_pole).float()
    idx = torch.multinomial(w, n_struct, replacement=True, generator=g)  # [B, ns]

    def gat(t):
        return torch.gather(t, 1, idx)

    k_s, ax_s, ay_s, bx_s, by_s, h_s = gat(kind), gat(ax), gat(ay), gat(bx), gat(by), gat(hh)
    u1 = torch.rand(B, n_struct, generator=g, device=dev, dtype=torch.float64)
    u2 = torch.rand(B, n_struct, generator=g, device=dev, dtype=torch.float64)
    u3 = torch.rand(B, n_struct, generator=g, device=dev, dtype=torch.float64)
    is_face = k_s == 1
    rr = bx_s * torch.sqrt(u1)
    px = torch.where(is_face, ax_s + u1 * (bx_s - ax_s), ax_s + rr * torch.cos(2 * math.pi * u2))
    py = torch.where(is_face, ay_s + u1 * (by_s - ay_s), ay_s + rr * torch.sin(2 * math.pi * u2))
    pz = GROUND_Z + u3 * h_s
    # world -> sensor frame
    c, s = torch.cos(th), torch.sin(th)
    dx, dy = px - sx, py - sy
    xs = c * dx + s * dy
    ys = -s * dx + c * dy

    ug = torch.rand(B, n_ground, generator=g, device=dev, dtype=torch.float64)
    az = torch.rand(B, n_ground, generator=g, device=dev, dtype=torch.float64) * (2 * math.pi)
    rg = 3 + 57 * ug ** 1.5
    xg, yg = rg * torch.cos(az), rg * torch.sin(az)
    zg = torch.full_like(xg, GROUND_Z)

    x = torch.cat([xg, xs], 1)
    y = torch.cat([yg, ys], 1)
    z = torch.cat([zg, pz], 1)
    noise = torch.randn(B, n_pts, 3, generator=g, device=dev, dtype=torch.float64) * 0.02
    x, y, z = x + noise[..., 0], y + noise[..., 1], z + noise[..., 2]
    order = torch.argsort(torch.atan2(y, x), dim=1)  # azimuth sweep order
    out = torch.zeros(B, n_pts, 4, dtype=torch.float32, device=dev)
    out[..., 0] = torch.gather(x, 1, order).float()
    out[..., 1] = torch.gather(y, 1, order).float()
    out[..., 2] = torch.gather(z, 1, order).float()
    return out


def db_layout(n_scans: int, visits_per_scene: int = 4, first_scene: int = 0):
    """(scene_seeds, visits) of a DB of n_scans = scenes x visits, scene-major."""
    seeds = [first_scene + i // visits_per_scene for i in range(n_scans)]
    visits = [i % visits_per_scene for i in range(n_scans)]
    return seeds, visits
