"""The reference's own tensor chain in front of the pose solve, restated with the same torch ops
on whatever device the inputs live on (the reference hard-codes .cuda()): flow2coord
(models/SFMnet.py:298-318), the point selection of pose_by_ransac (:239-254), bmm with K^-1
(:259-260), transpose/[:, :2]/contiguous (:262-263), .double() (epipolar_utils.py:130).
Test infrastructure — the plain-PyTorch reference of the flow_points kernel."""
import numpy as np
import torch
import torch.nn.functional as F


def flow2coord(flow):
    b, _, h, w = flow.size()
    coord1 = torch.zeros_like(flow)
    coord1[:, 0, :, :] += torch.arange(w, device=flow.device).float()
    coord1[:, 1, :, :] += torch.arange(h, device=flow.device).float()[:, None]
    coord2 = coord1 + flow
    ones = torch.ones((b, 1, h * w), dtype=torch.float32, device=flow.device)
    return (torch.cat((coord1.reshape(b, 2, h * w), ones), dim=1),
            torch.cat((coord2.reshape(b, 2, h * w), ones), dim=1))


def points_of_image(flow, Kinv, batch, margin=10, pts=None, sample_sp=False):
    """x1, x2 float64 [n,2] of image `batch`, as pose_by_ransac hands them to computeP."""
    b, _, h, w = flow.size()
    c1, c2 = flow2coord(flow)
    c1, c2 = c1.view(b, 3, h, w), c2.view(b, 3, h, w)
    if pts is None:
        a = c1[batch, :, margin:-margin, margin:-margin].contiguous().view(3, -1).unsqueeze(0)
        bb = c2[batch, :, margin:-margin, margin:-margin].contiguous().view(3, -1).unsqueeze(0)
    elif sample_sp:
        p = torch.from_numpy(np.asarray(pts)).to(c1.device).type_as(c1).clone()
        p[:, 0] = 2.0 * p[:, 0] / max(w - 1, 1) - 1.0
        p[:, 1] = 2.0 * p[:, 1] / max(h - 1, 1) - 1.0
        a = F.grid_sample(c1[batch].unsqueeze(0), p.unsqueeze(0).unsqueeze(-2), align_corners=True).squeeze(-1)
        bb = F.grid_sample(c2[batch].unsqueeze(0), p.unsqueeze(0).unsqueeze(-2), align_corners=True).squeeze(-1)
    else:
        p = np.int32(np.round(np.asarray(pts)))
        a = c1[batch, :, p[:, 1], p[:, 0]].unsqueeze(0)
        bb = c2[batch, :, p[:, 1], p[:, 0]].unsqueeze(0)
    Ki = Kinv[batch].unsqueeze(0)
    a = torch.bmm(Ki, a).transpose(1, 2)[0, :, :2].contiguous()
    bb = torch.bmm(Ki, bb).transpose(1, 2)[0, :, :2].contiguous()
    return a.double(), bb.double()
