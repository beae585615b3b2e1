"""Synthetic KITTI-shaped two-view scenes.

Pure numpy; used by the tests and by bench.py to make correspondences of the shape
`models/SFMnet.py:pose_by_ransac` feeds to `essential_matrix.computeP`
(reference: models/SFMnet.py:239-263 -- pixel grid + flow, K^-1 normalisation in float32,
then `.double()` in epipolar_utils.py:130).
"""
import numpy as np

KITTI_K = np.array([[721.5377, 0.0, 609.5593], [0.0, 721.5377, 172.854], [0.0, 0.0, 1.0]])
KITTI_HW = (370, 1226)


def rodrigues(r):
    r = np.asarray(r, dtype=np.float64)
    th = np.linalg.norm(r)
    if th < 1e-300:
        return np.eye(3)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * (Kx @ Kx)


def essential_from_pose(R, t):
    """E with x2^T E x1 = 0 for X2 = R X1 + t  (E = [t]x R)."""
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    return tx @ R


def make_pair(n=10000, seed=1234, rvec=(0.002, 0.01, -0.001), t=(0.03, -0.01, -0.8),
              noise_px=0.05, outlier_frac=0.2, outlier_px=30.0, dense=False, f32_origin=True,
              K=KITTI_K, hw=KITTI_HW, margin=10):
    """Returns dict(x1, x2 [N,2] float64 normalised coords, R, t (unit), E_gt, inlier_gt)."""
    rng = np.random.default_rng(seed)
    H, W = hw
    if dense:
        v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64),
                           indexing="ij")
        u, v = u.reshape(-1), v.reshape(-1)
        n = u.size
    else:
        u = rng.uniform(margin, W - margin, n)
        v = rng.uniform(margin, H - margin, n)
    depth = rng.uniform(5.0, 80.0, n)
    R = rodrigues(rvec)
    t = np.asarray(t, dtype=np.float64)
    Kinv = np.linalg.inv(K)
    p1 = np.stack([u, v, np.ones(n)], 0)
    X1 = (Kinv @ p1) * depth
    X2 = R @ X1 + t[:, None]
    p2 = K @ (X2 / X2[2:3])
    u2 = p2[0] + rng.normal(0.0, noise_px, n)
    v2 = p2[1] + rng.normal(0.0, noise_px, n)
    is_out = rng.uniform(0, 1, n) < outlier_frac
    u2 = u2 + is_out * rng.uniform(-outlier_px, outlier_px, n)
    v2 = v2 + is_out * rng.uniform(-outlier_px, outlier_px, n)
    if f32_origin:  # SFMnet normalises in float32 and up-casts
        K32 = Kinv.astype(np.float32)
        a = (K32 @ np.stack([u, v, np.ones(n)], 0).astype(np.float32))[:2].T
        b = (K32 @ np.stack([u2, v2, np.ones(n)], 0).astype(np.float32))[:2].T
        x1 = np.ascontiguousarray(a, dtype=np.float64)
        x2 = np.ascontiguousarray(b, dtype=np.float64)
    else:
        x1 = np.ascontiguousarray((Kinv @ np.stack([u, v, np.ones(n)], 0))[:2].T)
        x2 = np.ascontiguousarray((Kinv @ np.stack([u2, v2, np.ones(n)], 0))[:2].T)
    return dict(x1=x1, x2=x2, R=R, t=t / np.linalg.norm(t), E_gt=essential_from_pose(R, t),
                inlier_gt=~is_out, K=K)


def make_sets(n_points, n_sets, seed=1234):
    """Host-supplied minimal-set index table [H,5] int32 (throughput runs; parity runs use the
    reference's curand stream instead, see tv5.ref_rng_sets)."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, n_points, size=(n_sets, 5), dtype=np.int32)


def pair_variation(i):
    """Per-pair motion for batches (config 3): +-20 % variation of yaw and |t|."""
    rng = np.random.default_rng(99991 + i)
    s_yaw, s_t = rng.uniform(0.8, 1.2, 2)
    return dict(seed=1234 + i, rvec=(0.002, 0.01 * s_yaw, -0.001),
                t=(0.03 * s_t, -0.01 * s_t, -0.8 * s_t))


def rotation_error_deg(R_est, R_gt):
    c = (np.trace(R_est.T @ R_gt) - 1.0) / 2.0
    return float(np.degrees(np.arccos(np.clip(c, -1.0, 1.0))))


def translation_error_deg(t_est, t_gt):
    a = t_est / np.linalg.norm(t_est)
    b = t_gt / np.linalg.norm(t_gt)
    return float(np.degrees(np.arccos(np.clip(a @ b, -1.0, 1.0))))


def essential_distance(Ea, Eb):
    """Frobenius distance after normalisation and sign alignment."""
    a = Ea / np.linalg.norm(Ea)
    b = Eb / np.linalg.norm(Eb)
    return float(min(np.linalg.norm(a - b), np.linalg.norm(a + b)))


def make_flow(hw=KITTI_HW, seed=1234, rvec=(0.002, 0.01, -0.001), t=(0.03, -0.01, -0.8), noise_px=0.05,
              outlier_frac=0.2, outlier_px=30.0, K=KITTI_K):
    """Dense synthetic optical flow [2,H,W] float32 of a rigid scene with random depth (what the
    flow network of SFMnet would output, models/SFMnet.py:120-122), plus ground truth."""
    rng = np.random.default_rng(seed)
    H, W = hw
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    depth = rng.uniform(5.0, 80.0, (H, W))
    R = rodrigues(rvec)
    t = np.asarray(t, dtype=np.float64)
    Kinv = np.linalg.inv(K)
    p1 = np.stack([u, v, np.ones((H, W))], 0).reshape(3, -1)
    X2 = R @ ((Kinv @ p1) * depth.reshape(1, -1)) + t[:, None]
    p2 = (K @ (X2 / X2[2:3])).reshape(3, H, W)
    du = p2[0] - u + rng.normal(0.0, noise_px, (H, W))
    dv = p2[1] - v + rng.normal(0.0, noise_px, (H, W))
    is_out = rng.uniform(0, 1, (H, W)) < outlier_frac
    du = du + is_out * rng.uniform(-outlier_px, outlier_px, (H, W))
    dv = dv + is_out * rng.uniform(-outlier_px, outlier_px, (H, W))
    return dict(flow=np.stack([du, dv]).astype(np.float32), Kinv=Kinv.astype(np.float32), R=R,
                t=t / np.linalg.norm(t), E_gt=essential_from_pose(R, t))
