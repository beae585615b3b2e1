"""TEST / BENCH INFRASTRUCTURE.  Synthetic KITTI-shaped inputs for the reference's SFMnet.forward
(configs[4]): a textured image pair related by the optical flow of a rigid scene with smooth depth,
the intrinsics, and the ground-truth pose.  numpy + cv2 only."""
import numpy as np

from tv5 import synth


def smooth_depth(hw, seed):
    """Road-like scene: depth falls with the image row below the horizon, plus smooth bumps."""
    import cv2
    rng = np.random.default_rng(seed)
    H, W = hw
    v = np.arange(H, dtype=np.float64)[:, None] * np.ones((1, W))
    ground = 1.65 * 721.5 / np.maximum(v - 150.0, 3.0)            # camera 1.65 m above a plane
    bumps = cv2.GaussianBlur(rng.uniform(0.0, 1.0, (H, W)), (0, 0), 25.0)
    bumps = (bumps - bumps.min()) / (bumps.max() - bumps.min())
    return np.clip(np.minimum(ground, 8.0 + 60.0 * bumps), 4.0, 80.0)


def texture(hw, seed):
    """Band-limited noise in [-1, 1], 3 mostly luminance-correlated channels (SIFT finds several
    thousand keypoints, about 4,000 ratio-test matches: a KITTI-like count)."""
    import cv2
    rng = np.random.default_rng(seed)
    H, W = hw
    img = np.zeros((H, W, 3))
    for sigma in (1.5, 3.0, 6.0):
        n = rng.normal(0.0, 1.0, (H, W, 1)) * np.ones((1, 1, 3)) + 0.3 * rng.normal(0.0, 1.0, (H, W, 3))
        n = cv2.GaussianBlur(n, (0, 0), sigma)
        img += n / n.std()
    img = (img - img.min()) / (img.max() - img.min())
    return (2.0 * img - 1.0).astype(np.float32)


def make_scene(seed=0, hw=synth.KITTI_HW, rvec=(0.002, 0.01, -0.001), t=(0.03, -0.01, -0.8), noise_px=0.05):
    """dict(ref, target [3,H,W] f32 in [-1,1]; flow [2,H,W] f32 (ref pixel -> target pixel);
    K [3,3]; R, t (unit) with X_target = R X_ref + t)."""
    import cv2
    H, W = hw
    K = synth.KITTI_K
    rng = np.random.default_rng(seed + 17)
    depth = smooth_depth(hw, seed)
    R = synth.rodrigues(rvec)
    tv = np.asarray(t, dtype=np.float64)
    v, u = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    p1 = np.stack([u, v, np.ones((H, W))], 0).reshape(3, -1)
    X2 = R @ ((np.linalg.inv(K) @ p1) * depth.reshape(1, -1)) + tv[:, None]
    p2 = (K @ (X2 / X2[2:3])).reshape(3, H, W)
    du, dv = p2[0] - u, p2[1] - v
    flow = np.stack([du + rng.normal(0, noise_px, (H, W)), dv + rng.normal(0, noise_px, (H, W))]).astype(np.float32)
    ref = texture(hw, seed)
    # target(x + flow(x)) = ref(x): backward warp with the (smooth) flow as its own inverse estimate
    mapx = (u - du).astype(np.float32)
    mapy = (v - dv).astype(np.float32)
    tgt = cv2.remap(ref, mapx, mapy, cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT)
    return dict(ref=np.ascontiguousarray(ref.transpose(2, 0, 1)), target=np.ascontiguousarray(tgt.transpose(2, 0, 1)),
                flow=flow, K=K.copy(), R=R, t=tv / np.linalg.norm(tv), depth=depth)
