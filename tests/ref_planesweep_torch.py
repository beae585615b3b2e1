"""Plain-PyTorch reference of the plane_sweep kernel: the torch operator sequence the reference
executes for its plane-sweep volume — PSNet.forward's label loop (models/PSNet.py:141-157) calling
inverse_warp (models/inverse_warp.py:121-153; back-projection :31-45, projection and [-1,1]
normalisation with out-of-range -> 2 :48-78, grid_sample zeros / align_corners=True) — issued
op for op in the same order and dtype, so rounding is what the reference's would be on the device
the inputs live on (the reference itself hard-codes .cuda()).  Test infrastructure only."""
import torch
import torch.nn.functional as F


def _homogeneous_pixel_grid(h, w, like):
    ys, xs = torch.meshgrid(torch.arange(h, device=like.device), torch.arange(w, device=like.device), indexing="ij")
    return torch.stack((xs, ys, torch.ones_like(xs)), dim=0).type_as(like)          # [3, h, w], exact integers


def _sample_grid(depth, pose, K, K_inv):
    """[B, h, w, 2] normalised source coordinates of every target pixel at the given depth map."""
    b, h, w = depth.shape
    pix = _homogeneous_pixel_grid(h, w, depth).unsqueeze(0).expand(b, 3, h, w).contiguous().view(b, 3, -1)
    rays = K_inv.bmm(pix).view(b, 3, h, w)
    cam = rays * depth.unsqueeze(1)                                  # back-projected points
    proj = K.bmm(pose)                                               # [B, 3, 4]
    pc = proj[:, :, :3].bmm(cam.view(b, 3, -1)) + proj[:, :, -1:]
    z = pc[:, 2].clamp(min=1e-3)
    xn = 2 * (pc[:, 0] / z) / (w - 1) - 1
    yn = 2 * (pc[:, 1] / z) / (h - 1) - 1
    xn[((xn > 1) + (xn < -1)).detach()] = 2                          # zeros padding: push outside
    yn[((yn > 1) + (yn < -1)).detach()] = 2
    return torch.stack([xn, yn], dim=2).view(b, h, w, 2)


def warp_to_plane(feat, depth, pose, K, K_inv):
    return F.grid_sample(feat, _sample_grid(depth, pose, K, K_inv), padding_mode="zeros", align_corners=True)


def cost_volume(ref_fea, tgt_fea, pose, intrinsics4, intrinsics_inv4, nlabel, mindepth, by_depth=False):
    b, c, h, w = ref_fea.shape
    unit = torch.ones(b, h, w, device=ref_fea.device)
    inverse_depth_scale = unit * mindepth * nlabel
    volume = torch.zeros(b, 2 * c, int(nlabel), h, w, device=ref_fea.device)
    for i in range(int(nlabel)):
        depth = unit * (i + 1) * mindepth if by_depth else torch.div(inverse_depth_scale, i + 1 + 1e-16)
        volume[:, :c, i] = ref_fea
        volume[:, c:, i] = warp_to_plane(tgt_fea, depth, pose, intrinsics4, intrinsics_inv4)
    return volume.contiguous()
