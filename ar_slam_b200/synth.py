"""Synthetic maps of the BASELINE shapes.

Tags lie in a planar-ish slab on a jittered grid (0.3 m pitch, the demo's
spacing), each rotated by up to 30 degrees off the common normal; tag side
0.0635 m (reference ar_slam_util.hpp:319).  Captures look at the slab from
1 - 2.5 m with up to 30 degrees of tilt and a random roll; f = 760 px, image
1020 x 768 (the demo's values).  Every capture observes the `tags_per_capture`
tags nearest to its look-at point that project inside the image; observations
are the exact projection + N(0, noise_px) and go through a float32 round trip
like geometry_msgs/Point32 (ar_slam_interfaces/msg/Detection.msg).
Everything derives from one Philox stream keyed by `seed`, so every rank can
regenerate the same map.

The projection below is plain numpy and only generates data; it is not a
solver path.
"""
import numpy as np
from scipy.spatial import cKDTree

TAG_SIZE = 0.0635
IMG_W, IMG_H = 1020, 768
F_TRUE = 760.0
_DIRS = np.array([[-1.0, -1.0], [1.0, -1.0], [1.0, 1.0], [-1.0, 1.0]])


def rodrigues(aa):
    """Rotation matrices [n,3,3] of angle-axis vectors [n,3]."""
    aa = np.asarray(aa, dtype=np.float64).reshape(-1, 3)
    th = np.linalg.norm(aa, axis=1)
    safe = np.where(th > 1e-12, th, 1.0)
    k = aa / safe[:, None]
    K = np.zeros((len(aa), 3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    s, c = np.sin(th)[:, None, None], np.cos(th)[:, None, None]
    R = np.eye(3)[None] + s * K + (1 - c) * (K @ K)
    R[th <= 1e-12] = np.eye(3)
    return R


def matrix_to_aa(R):
    """Angle-axis vectors [n,3] of rotation matrices [n,3,3] (angles < pi)."""
    R = np.asarray(R).reshape(-1, 3, 3)
    c = np.clip((np.trace(R, axis1=1, axis2=2) - 1) / 2, -1, 1)
    th = np.arccos(c)
    v = np.stack([R[:, 2, 1] - R[:, 1, 2], R[:, 0, 2] - R[:, 2, 0], R[:, 1, 0] - R[:, 0, 1]], axis=1)
    s = 2 * np.sin(th)
    out = np.where(s[:, None] > 1e-12, v / np.where(s > 1e-12, s, 1.0)[:, None] * th[:, None], v / 2)
    return out


def project(cam, cap_pose, tag_pose, tag_size=TAG_SIZE):
    """Corner pixels [n,4,2] and depths [n,4] of n (capture, tag) pairs.  cam = (f, l1, l2): with
    l1 = l2 = 0 the reference's live focal-only model, otherwise the radial model of the TODO at
    ar_slam_util.cpp:164-171, f (1 + l1 r^2 + l2 r^4) (x', y')."""
    Ra, Rc = rodrigues(tag_pose[:, 3:]), rodrigues(cap_pose[:, 3:])
    m = np.zeros((4, 3))
    m[:, :2] = 0.5 * tag_size * _DIRS
    pw = np.einsum("nij,kj->nki", Ra, m) + tag_pose[:, None, :3]
    q = pw + cap_pose[:, None, :3]
    p = np.einsum("nij,nkj->nki", Rc, q)
    xy = p[..., :2] / p[..., 2:3]
    if len(cam) >= 3 and (cam[1] != 0.0 or cam[2] != 0.0):
        r2 = (xy * xy).sum(-1, keepdims=True)
        xy = xy * (r2 * (cam[1] + cam[2] * r2) + 1.0)
    uv = cam[0] * xy
    return uv, p[..., 2]


class SynthMap:
    pass


def make_map(n_cap, n_tag, tags_per_capture=8, seed=0xA55A0002, noise_px=0.3, pitch=0.3,
             init_noise_t=0.02, init_noise_r_deg=2.0, f_init=800.0, distortion=(0.0, 0.0)):
    """Returns a SynthMap with ground truth, observations and a perturbed initial state."""
    rng = np.random.Generator(np.random.Philox(key=int(seed)))
    # ---- tags on a jittered grid
    side = int(np.ceil(np.sqrt(n_tag)))
    gi, gj = np.meshgrid(np.arange(side), np.arange(side), indexing="ij")
    grid = np.stack([gi.ravel(), gj.ravel()], axis=1)[:n_tag].astype(np.float64)
    tag_t = np.zeros((n_tag, 3))
    tag_t[:, :2] = grid * pitch + rng.uniform(-0.3 * pitch, 0.3 * pitch, (n_tag, 2))
    tag_t[:, 2] = rng.uniform(-0.1, 0.1, n_tag)
    axis = rng.normal(size=(n_tag, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    tag_w = axis * rng.uniform(0, np.deg2rad(30.0), (n_tag, 1))
    tag_pose = np.concatenate([tag_t, tag_w], axis=1)
    extent = side * pitch
    tree = cKDTree(tag_t[:, :2])
    cam_true = np.array([F_TRUE, float(distortion[0]), float(distortion[1])])

    cap_pose = np.zeros((n_cap, 6))
    blk_cap, blk_tag, blk_obs = [], [], []
    todo = np.arange(n_cap)
    kq = min(n_tag, max(3 * tags_per_capture, 24))
    for _ in range(200):
        if len(todo) == 0:
            break
        n = len(todo)
        look = np.zeros((n, 3))
        look[:, :2] = rng.uniform(-0.02 * extent, 1.0 * extent, (n, 2))
        h = rng.uniform(1.0, 2.5, n)
        tilt_axis = rng.normal(size=(n, 3))
        tilt_axis[:, 2] = 0
        tilt_axis /= np.linalg.norm(tilt_axis, axis=1, keepdims=True)
        tilt = rodrigues(tilt_axis * rng.uniform(0, np.deg2rad(30.0), (n, 1)))
        roll = rodrigues(np.stack([np.zeros(n), np.zeros(n), rng.uniform(-np.pi, np.pi, n)], axis=1))
        Rc = roll @ tilt  # world -> camera
        # camera centre: back off from the look-at point along the optical axis
        zc = Rc[:, 2, :]  # optical axis in world coordinates
        centre = look - zc * h[:, None]
        pose = np.concatenate([-centre, matrix_to_aa(Rc)], axis=1)
        _, nn = tree.query(look[:, :2], k=kq)
        nn = nn.reshape(n, kq)
        uv, z = project(cam_true, np.repeat(pose, kq, axis=0), tag_pose[nn.ravel()])
        uv, z = uv.reshape(n, kq, 4, 2), z.reshape(n, kq, 4)
        inside = (np.abs(uv[..., 0]) < 0.5 * IMG_W - 4).all(-1) & (np.abs(uv[..., 1]) < 0.5 * IMG_H - 4).all(-1) \
            & (z > 0.2).all(-1)
        rank = np.cumsum(inside, axis=1)
        take = inside & (rank <= tags_per_capture)
        ok = take.sum(1) >= tags_per_capture
        rows, cols = np.nonzero(take & ok[:, None])   # row-major: grouped by capture, nearest first
        cap_pose[todo[ok]] = pose[ok]
        blk_cap.append(todo[rows])
        blk_tag.append(nn[rows, cols])
        blk_obs.append(uv[rows, cols].reshape(len(rows), 8))
        todo = todo[~ok]
    if len(todo):
        raise RuntimeError("could not place %d captures" % len(todo))
    blk_cap = np.concatenate(blk_cap).astype(np.int32)
    order = np.argsort(blk_cap, kind="stable")
    blk_cap = blk_cap[order]
    blk_tag = np.concatenate(blk_tag).astype(np.int32)[order]
    obs = np.concatenate(blk_obs)[order]
    obs = obs + rng.normal(0, noise_px, obs.shape)
    obs = obs.astype(np.float32).astype(np.float64)  # Point32 round trip

    m = SynthMap()
    m.n_cap, m.n_tag = n_cap, n_tag
    m.cap_idx, m.tag_idx, m.obs = blk_cap, blk_tag, np.ascontiguousarray(obs)
    m.cam_true, m.cap_true, m.tag_true = cam_true, cap_pose, tag_pose
    # ---- initial state: ground truth perturbed by init_noise_t metres / init_noise_r_deg degrees
    m.cam0 = np.array([f_init, 0.0, 0.0])
    m.cap0 = cap_pose + np.concatenate([rng.normal(0, init_noise_t, (n_cap, 3)),
                                        rng.normal(0, np.deg2rad(init_noise_r_deg), (n_cap, 3))], axis=1)
    m.tag0 = tag_pose + np.concatenate([rng.normal(0, init_noise_t, (n_tag, 3)),
                                        rng.normal(0, np.deg2rad(init_noise_r_deg), (n_tag, 3))], axis=1)
    used = np.zeros(n_tag, bool)
    used[blk_tag] = True
    m.tags_used = int(used.sum())
    return m


def make_localization_batch(n_loc, n_tag, tags_per_capture=8, seed=0xA55A0004, noise_px=0.3):
    """Config 4: captures to localise against the fixed ground-truth map."""
    m = make_map(n_loc, n_tag, tags_per_capture, seed=seed, noise_px=noise_px)
    counts = np.bincount(m.cap_idx, minlength=n_loc)
    m.blk_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    m.seed_block = np.zeros(n_loc, dtype=np.int32)  # first block: every tag is in the map
    return m


def write_detections_yaml(m, path, width=IMG_W, height=IMG_H, f0=3000.0):
    """The synthetic map's observations in the reference's detections / map.yaml layout
    (ar_slam_util.cpp:304-465: blocks, captures, arucos, camera), poses at zero and f at the reference's
    initial 3000 (ar_slam_util.hpp:69), so that ar_slam_cli builds the map from scratch with the
    reference's own schedule and seed heuristics."""
    with open(path, "w") as f:
        f.write("blocks:\n")
        for c, t, r in zip(m.cap_idx, m.tag_idx, m.obs):
            f.write("  - capture: cap_%d\n    aruco: aruco_4X4_50_%d\n    aruco_rect: [%s]\n"
                    % (c, t, ", ".join(repr(float(v)) for v in r)))
        f.write("captures:\n")
        for c in range(m.n_cap):
            f.write("  cap_%d:\n    inv_pose: [0, 0, 0, 0, 0, 0]\n    img_fn: cap_%d.jpg\n" % (c, c))
        f.write("arucos:\n")
        for t in range(m.n_tag):
            f.write("  aruco_4X4_50_%d:\n    pose: [0, 0, 0, 0, 0, 0]\n" % t)
        f.write("camera:\n  params: [%r, 0, 0]\n  width: %d\n  height: %d\n" % (float(f0), width, height))


# ---- camera frames for the marker detector ------------------------------------------
def dict_4x4_50_bits():
    """(50, 4, 4) uint8 cells of cv::aruco::DICT_4X4_50, read from the table the library embeds
    (csrc/dict_4x4_50.inc, written by scripts/make_dictionary_table.py)."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "dict_4x4_50.inc")
    codes = [int(c, 16) for c in re.findall(r"0x([0-9a-fA-F]{4})", open(path).read())]
    assert len(codes) == 50
    return np.array([[[(c >> (15 - (4 * y + x))) & 1 for x in range(4)] for y in range(4)] for c in codes], np.uint8)


def render_marker_scene(h, w, dict_bits, n_markers, seed, noise=4.0, blur=True):
    """Synthetic camera frame: dictionary markers (black border, white quiet zone) under random homographies on a
    shaded background, box-blurred, with Gaussian pixel noise.  Pure numpy (no OpenCV), deterministic per seed.
    Returns (bgr uint8 image (h, w, 3), list of the marker ids placed)."""
    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = 120 + 50 * np.sin(xx / w * 3.1 + rng.uniform(0, 6)) * np.cos(yy / h * 2.3 + rng.uniform(0, 6))
    ids = []
    placed = []
    for _ in range(n_markers * 6):
        if len(ids) >= n_markers:
            break
        side = rng.uniform(0.05, 0.16) * w
        cx, cy = rng.uniform(side, w - side), rng.uniform(side, h - side)
        if any((cx - px) ** 2 + (cy - py) ** 2 < (1.1 * (side + ps)) ** 2 for px, py, ps in placed):
            continue
        placed.append((cx, cy, side))
        mid = int(rng.integers(0, dict_bits.shape[0]))
        ids.append(mid)
        ang = rng.uniform(0, 2 * np.pi)
        tilt = rng.uniform(0.6, 1.0)
        persp = rng.uniform(-0.15, 0.15, 2) / side
        c, s = np.cos(ang), np.sin(ang)
        # marker plane coordinates (u, v) in [-1, 1] cover the marker with its border; quiet zone out to 1.35
        a = np.array([[c * side / 2, -s * side / 2 * tilt, cx], [s * side / 2, c * side / 2 * tilt, cy],
                      [persp[0], persp[1], 1.0]])
        inv = np.linalg.inv(a)
        r = int(side * 1.2) + 2
        x0, x1 = max(0, int(cx) - r), min(w, int(cx) + r)
        y0, y1 = max(0, int(cy) - r), min(h, int(cy) + r)
        sub_y, sub_x = np.mgrid[y0:y1, x0:x1].astype(np.float64)
        den = inv[2, 0] * sub_x + inv[2, 1] * sub_y + inv[2, 2]
        u = (inv[0, 0] * sub_x + inv[0, 1] * sub_y + inv[0, 2]) / den
        v = (inv[1, 0] * sub_x + inv[1, 1] * sub_y + inv[1, 2]) / den
        n = dict_bits.shape[1] + 2
        cell_x = np.floor((u + 1) / 2 * n).astype(int)
        cell_y = np.floor((v + 1) / 2 * n).astype(int)
        full = np.zeros((n, n))
        full[1:-1, 1:-1] = dict_bits[mid]
        inside = (cell_x >= 0) & (cell_x < n) & (cell_y >= 0) & (cell_y < n)
        quiet = (np.abs(u) < 1.35) & (np.abs(v) < 1.35)
        val = np.where(inside, full[np.clip(cell_y, 0, n - 1), np.clip(cell_x, 0, n - 1)] * 200 + 30, 235.0)
        region = img[y0:y1, x0:x1]
        region[quiet] = val[quiet]
    if blur:
        pad = np.pad(img, 1, mode="edge")
        img = (pad[:-2, 1:-1] + 2 * pad[1:-1, 1:-1] + pad[2:, 1:-1]) / 4
        pad = np.pad(img, 1, mode="edge")
        img = (pad[1:-1, :-2] + 2 * pad[1:-1, 1:-1] + pad[1:-1, 2:]) / 4
    img = img + rng.normal(0, noise, img.shape)
    g = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    bgr = np.stack([np.clip(g.astype(int) + d, 0, 255).astype(np.uint8) for d in (-3, 0, 4)], axis=-1)
    return bgr, ids
