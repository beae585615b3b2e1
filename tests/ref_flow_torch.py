"""Plain-PyTorch reference of the flow_points kernel: the torch operator sequence the reference
executes in front of its pose solve — flow2coord (models/SFMnet.py:298-318), the point selection of
pose_by_ransac (:239-254), bmm with K^-1 (:259-260), transpose/[:, :2]/contiguous (:262-263) and
.double() (epipolar_utils.py:130) — issued op for op in the same order and dtype on whatever device
the inputs live on (the reference hard-codes .cuda()).  Test infrastructure only."""
import numpy as np
import torch
import torch.nn.functional as F


def _homogeneous_coords(flow):
    """(pixel grid, pixel grid + flow), both [B, 3, h, w] float32 with a row of ones."""
    b, _, h, w = flow.shape
    grid = torch.zeros_like(flow)
    grid[:, 0] += torch.arange(w, device=flow.device).float()
    grid[:, 1] += torch.arange(h, device=flow.device).float()[:, None]
    moved = grid + flow
    ones = torch.ones((b, 1, h * w), dtype=torch.float32, device=flow.device)
    c1 = torch.cat((grid.reshape(b, 2, h * w), ones), dim=1).view(b, 3, h, w)
    c2 = torch.cat((moved.reshape(b, 2, h * w), ones), dim=1).view(b, 3, h, w)
    return c1, c2


def points_of_image(flow, Kinv, batch, margin=10, pts=None, sample_sp=False):
    """x1, x2 float64 [n,2] of image `batch`, as pose_by_ransac hands them to computeP."""
    b, _, h, w = flow.shape
    c1, c2 = _homogeneous_coords(flow)
    if pts is None:                                   # dense crop
        sel = [c[batch, :, margin:-margin, margin:-margin].contiguous().view(3, -1).unsqueeze(0) for c in (c1, c2)]
    elif sample_sp:                                   # bilinear at sub-pixel keypoints
        g = torch.from_numpy(np.asarray(pts)).to(c1.device).type_as(c1).clone()
        g[:, 0] = 2.0 * g[:, 0] / max(w - 1, 1) - 1.0
        g[:, 1] = 2.0 * g[:, 1] / max(h - 1, 1) - 1.0
        g = g.unsqueeze(0).unsqueeze(-2)
        sel = [F.grid_sample(c[batch].unsqueeze(0), g, align_corners=True).squeeze(-1) for c in (c1, c2)]
    else:                                             # rounded keypoints
        k = np.int32(np.round(np.asarray(pts)))
        sel = [c[batch, :, k[:, 1], k[:, 0]].unsqueeze(0) for c in (c1, c2)]
    Ki = Kinv[batch].unsqueeze(0)
    out = [torch.bmm(Ki, s).transpose(1, 2)[0, :, :2].contiguous().double() for s in sel]
    return out[0], out[1]
