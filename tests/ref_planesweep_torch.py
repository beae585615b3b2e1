"""The reference's plane-sweep cost-volume assembly restated with the same torch ops on whatever
device the inputs live on (the reference hard-codes .cuda()): PSNet.forward's label loop
(models/PSNet.py:141-157) over inverse_warp (models/inverse_warp.py:121-153: pixel2cam :31-45,
cam2pixel :48-78, grid_sample zeros/align_corners=True).  Test infrastructure — the plain-PyTorch
reference of the plane_sweep kernel."""
import torch
import torch.nn.functional as F


def pixel2cam(depth, intrinsics_inv):
    b, h, w = depth.size()
    i_range = torch.arange(0, h, device=depth.device).view(1, h, 1).expand(1, h, w).type_as(depth)
    j_range = torch.arange(0, w, device=depth.device).view(1, 1, w).expand(1, h, w).type_as(depth)
    ones = torch.ones(1, h, w, device=depth.device).type_as(depth)
    pixel_coords = torch.stack((j_range, i_range, ones), dim=1)
    cur = pixel_coords[:, :, :h, :w].expand(b, 3, h, w).contiguous().view(b, 3, -1)
    cam_coords = intrinsics_inv.bmm(cur).view(b, 3, h, w)
    return cam_coords * depth.unsqueeze(1)


def cam2pixel(cam_coords, proj_c2p_rot, proj_c2p_tr, padding_mode):
    b, _, h, w = cam_coords.size()
    flat = cam_coords.view(b, 3, -1)
    pcoords = proj_c2p_rot.bmm(flat) + proj_c2p_tr
    X, Y = pcoords[:, 0], pcoords[:, 1]
    Z = pcoords[:, 2].clamp(min=1e-3)
    X_norm = 2 * (X / Z) / (w - 1) - 1
    Y_norm = 2 * (Y / Z) / (h - 1) - 1
    if padding_mode == 'zeros':
        X_mask = ((X_norm > 1) + (X_norm < -1)).detach()
        X_norm[X_mask] = 2
        Y_mask = ((Y_norm > 1) + (Y_norm < -1)).detach()
        Y_norm[Y_mask] = 2
    return torch.stack([X_norm, Y_norm], dim=2).view(b, h, w, 2)


def inverse_warp(feat, depth, pose, intrinsics, intrinsics_inv, padding_mode='zeros'):
    cam_coords = pixel2cam(depth, intrinsics_inv)
    proj = intrinsics.bmm(pose)
    src = cam2pixel(cam_coords, proj[:, :, :3], proj[:, :, -1:], padding_mode)
    return F.grid_sample(feat, src, padding_mode=padding_mode, align_corners=True)


def cost_volume(ref_fea, tgt_fea, pose, intrinsics4, intrinsics_inv4, nlabel, mindepth, by_depth=False):
    b, c, h, w = ref_fea.shape
    ones_vec = torch.ones(b, h, w, device=ref_fea.device)
    disp2depth = ones_vec * mindepth * nlabel
    cost = torch.zeros(b, 2 * c, int(nlabel), h, w, device=ref_fea.device)
    for i in range(int(nlabel)):
        depth = ones_vec * (i + 1) * mindepth if by_depth else torch.div(disp2depth, i + 1 + 1e-16)
        warped = inverse_warp(tgt_fea, depth, pose, intrinsics4, intrinsics_inv4)
        cost[:, :c, i, :, :] = ref_fea
        cost[:, c:, i, :, :] = warped
    return cost.contiguous()
