"""Engine: one libtv5 context per CUDA device + tensor marshalling.

Mirrors the argument meaning of the reference extension
(RANSAC_FiveP/essential_matrix/essential_matrix_wrapper.cpp:45-108) and adds the batched,
mask-returning and hypothesis-table entry points of include/tv5.h.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import lib as _lib
from .lib import Tv5Error, Tv5Result

REF_THREADS = 512


def _check(ctx, rc, what):
    if rc != 0:
        L = _lib.load_library()
        msg = L.tv5_strerror(rc).decode()
        if rc == -2 and ctx is not None:
            msg += f" (cudaError {L.tv5_last_cuda_error(ctx)})"
        raise Tv5Error(f"{what}: {msg}")


def _require_points(x, name):
    # same checks and wording as CHECK_INPUT_INIT, essential_matrix_wrapper.cpp:39-42
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if x.dtype != torch.float64:
        raise RuntimeError(f"{name} must be a double tensor")
    if not x.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if x.dim() != 2 or x.shape[1] != 2:
        raise RuntimeError(f"{name} must have shape [N, 2]")


@dataclass
class PoseResult:
    """Outputs of one pose solve.  Tensors live on the device; `stats` is a device int32[8]
    (tv5_result) that is only read (with a stream sync) when one of the properties is used."""
    E: torch.Tensor
    P: torch.Tensor
    stats: torch.Tensor
    mask: torch.Tensor = None
    _host: np.ndarray = None

    def _h(self):
        if self._host is None:
            self._host = self.stats.cpu().numpy()
        return self._host

    @property
    def count(self):
        return int(self._h()[..., 0]) if self.stats.dim() == 1 else self._h()[:, 0].copy()

    @property
    def best_set(self):
        return int(self._h()[..., 1]) if self.stats.dim() == 1 else self._h()[:, 1].copy()

    @property
    def best_root(self):
        return int(self._h()[..., 2]) if self.stats.dim() == 1 else self._h()[:, 2].copy()

    @property
    def n_hypotheses(self):
        return int(self._h()[..., 3]) if self.stats.dim() == 1 else self._h()[:, 3].copy()

    @property
    def n_candidates(self):
        return int(self._h()[..., 4]) if self.stats.dim() == 1 else self._h()[:, 4].copy()

    @property
    def fast_path(self):
        return int(self._h()[..., 5]) if self.stats.dim() == 1 else self._h()[:, 5].copy()


class Engine:
    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise Tv5Error("tv5 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.L = _lib.load_library()
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None
                                   else torch.device(device).index or 0)
        h = C.c_void_p()
        _check(None, self.L.tv5_create(self.device.index, C.byref(h)), "tv5_create")
        self.ctx = h
        self.sm_count = self.L.tv5_device_sm_count(self.ctx)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.tv5_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- pose ---------------------------------------------------------------------------------
    def compute_pose(self, x1, x2, iters, thr, n_pre=None, n_full=None, sets=None,
                     with_cheirality=True, want_mask=False):
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        N = x1.shape[0]
        if x2.shape[0] != N or N < 1:
            raise RuntimeError("input1 and input2 must have the same, non-zero number of points")
        n_pre = N if n_pre is None else int(n_pre)
        n_full = N if n_full is None else int(n_full)
        with torch.cuda.device(self.device):
            E = torch.empty(3, 3, dtype=torch.float64, device=self.device)
            P = torch.empty(3, 4, dtype=torch.float64, device=self.device)
            stats = torch.empty(8, dtype=torch.int32, device=self.device)
            mask = torch.empty(min(n_full, N), dtype=torch.uint8, device=self.device) if want_mask else None
            sp = None
            if sets is not None:
                sets = self._sets(sets, iters)
                sp = sets.data_ptr()
            rc = self.L.tv5_compute_pose(self.ctx, self._stream(), x1.data_ptr(), x2.data_ptr(), N,
                                         sp, int(iters), n_pre, n_full, float(thr),
                                         int(bool(with_cheirality)), E.data_ptr(), P.data_ptr(),
                                         stats.data_ptr(), mask.data_ptr() if want_mask else None)
        _check(self.ctx, rc, "tv5_compute_pose")
        return PoseResult(E, P, stats, mask)

    def _sets(self, sets, iters, B=1):
        if not isinstance(sets, torch.Tensor):
            sets = torch.as_tensor(np.ascontiguousarray(sets, dtype=np.int32))
        sets = sets.to(device=self.device, dtype=torch.int32).contiguous()
        if sets.numel() != B * REF_THREADS * int(iters) * 5:
            raise RuntimeError(f"sets must hold {B}x{REF_THREADS * int(iters)}x5 indices")
        return sets

    def compute_pose_batch(self, x1, x2, offsets, iters, thr, n_pre=0, n_full=0, sets=None,
                           with_cheirality=True, want_mask=False):
        """x1, x2: [sum N_b, 2] float64 CUDA; offsets: [B+1] prefix sums (host ints)."""
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.size - 1
        if B < 1 or off[0] != 0 or off[-1] != x1.shape[0] or x2.shape[0] != x1.shape[0]:
            raise RuntimeError("offsets must be prefix sums covering all points")
        with torch.cuda.device(self.device):
            E = torch.empty(B, 3, 3, dtype=torch.float64, device=self.device)
            P = torch.empty(B, 3, 4, dtype=torch.float64, device=self.device)
            stats = torch.empty(B, 8, dtype=torch.int32, device=self.device)
            mask = torch.empty(x1.shape[0], dtype=torch.uint8, device=self.device) if want_mask else None
            sp = None
            if sets is not None:
                sets = self._sets(sets, iters, B)
                sp = sets.data_ptr()
            rc = self.L.tv5_compute_pose_batch(
                self.ctx, self._stream(), B, x1.data_ptr(), x2.data_ptr(),
                off.ctypes.data_as(C.POINTER(C.c_int64)), sp, int(iters), int(n_pre), int(n_full),
                float(thr), int(bool(with_cheirality)), E.data_ptr(), P.data_ptr(), stats.data_ptr(),
                mask.data_ptr() if want_mask else None)
        _check(self.ctx, rc, "tv5_compute_pose_batch")
        return PoseResult(E, P, stats, mask)

    def compute_pose_batch_host(self, x1, x2, offsets, iters, thr, n_pre=0, n_full=0, sets=None,
                                with_cheirality=True):
        """Host numpy buffers in, host numpy out (copies inside); synchronous."""
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        x2 = np.ascontiguousarray(x2, dtype=np.float64)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.size - 1
        E = np.empty((B, 3, 3))
        P = np.empty((B, 3, 4))
        res = (Tv5Result * B)()
        sp = None
        if sets is not None:
            sets = np.ascontiguousarray(sets, dtype=np.int32)
            sp = sets.ctypes.data
        with torch.cuda.device(self.device):
            rc = self.L.tv5_compute_pose_batch_host(
                self.ctx, self._stream(), B, x1.ctypes.data, x2.ctypes.data,
                off.ctypes.data_as(C.POINTER(C.c_int64)), sp, int(iters), int(n_pre), int(n_full),
                float(thr), int(bool(with_cheirality)), E.ctypes.data, P.ctypes.data,
                C.cast(res, C.c_void_p))
        _check(self.ctx, rc, "tv5_compute_pose_batch_host")
        stats = np.frombuffer(res, dtype=np.int32).reshape(B, 8).copy()
        return E, P, stats

    # -- building blocks ----------------------------------------------------------------------
    def solve5(self, x1, x2, sets, with_cheirality=True):
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        if not isinstance(sets, torch.Tensor):
            sets = torch.as_tensor(np.ascontiguousarray(sets, dtype=np.int32))
        sets = sets.to(device=self.device, dtype=torch.int32).contiguous().view(-1, 5)
        H = sets.shape[0]
        with torch.cuda.device(self.device):
            E = torch.empty(H, 10, 3, 3, dtype=torch.float64, device=self.device)
            P = torch.empty(H, 10, 3, 4, dtype=torch.float64, device=self.device)
            nr = torch.empty(H, dtype=torch.int32, device=self.device)
            nv = torch.empty(H, dtype=torch.int32, device=self.device)
            rc = self.L.tv5_solve5(self.ctx, self._stream(), x1.data_ptr(), x2.data_ptr(),
                                   x1.shape[0], sets.data_ptr(), H, int(bool(with_cheirality)),
                                   E.data_ptr(), P.data_ptr(), nr.data_ptr(), nv.data_ptr())
        _check(self.ctx, rc, "tv5_solve5")
        return dict(E=E, P=P, n_roots=nr, n_valid=nv)

    def score(self, x1, x2, E_list, thr, n_test=None, want_mask=False):
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        E_list = E_list.to(device=self.device, dtype=torch.float64).contiguous().view(-1, 9)
        M = E_list.shape[0]
        n = x1.shape[0] if n_test is None else int(n_test)
        with torch.cuda.device(self.device):
            counts = torch.empty(M, dtype=torch.int32, device=self.device)
            masks = torch.zeros(M, (n + 31) // 32, dtype=torch.int32, device=self.device) if want_mask else None
            rc = self.L.tv5_score(self.ctx, self._stream(), x1.data_ptr(), x2.data_ptr(), n,
                                  E_list.data_ptr(), M, float(thr), counts.data_ptr(),
                                  masks.data_ptr() if want_mask else None)
        _check(self.ctx, rc, "tv5_score")
        return (counts, masks) if want_mask else counts

    def score_bounds(self, x1, x2, E_list, thr, n_test=None):
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        E_list = E_list.to(device=self.device, dtype=torch.float64).contiguous().view(-1, 9)
        M = E_list.shape[0]
        n = x1.shape[0] if n_test is None else int(n_test)
        with torch.cuda.device(self.device):
            lo = torch.empty(M, dtype=torch.int32, device=self.device)
            hi = torch.empty(M, dtype=torch.int32, device=self.device)
            rc = self.L.tv5_score_bounds(self.ctx, self._stream(), x1.data_ptr(), x2.data_ptr(), n,
                                         E_list.data_ptr(), M, float(thr), lo.data_ptr(), hi.data_ptr())
        _check(self.ctx, rc, "tv5_score_bounds")
        return lo, hi

    def ref_rng_sets(self, N, iters):
        with torch.cuda.device(self.device):
            out = torch.empty(REF_THREADS * int(iters), 5, dtype=torch.int32, device=self.device)
            rc = self.L.tv5_ref_rng_sets(self.ctx, self._stream(), int(N), int(iters), out.data_ptr())
        _check(self.ctx, rc, "tv5_ref_rng_sets")
        return out

    # -- optical flow -> correspondences (front of pose_by_ransac, models/SFMnet.py:176-272) ---
    def _flow_args(self, flow, Kinv, margin, pts, offsets):
        if not isinstance(flow, torch.Tensor) or not flow.is_cuda or flow.dtype != torch.float32:
            raise RuntimeError("flow must be a CUDA float tensor")
        if flow.dim() != 4 or flow.shape[1] != 2:
            raise RuntimeError("flow must have shape [B, 2, H, W]")
        flow = flow.contiguous()
        B, _, H, W = flow.shape
        Kinv = Kinv.to(device=self.device, dtype=torch.float32).contiguous()
        if Kinv.numel() != B * 9:
            raise RuntimeError("intrinsic_inv must have shape [B, 3, 3]")
        if pts is None:
            mode, off, pp = 0, None, None
            n = (H - 2 * margin) * (W - 2 * margin)
            if margin < 0 or 2 * margin >= min(H, W):
                raise RuntimeError("margin leaves no pixels")
            counts = np.full(B, n, dtype=np.int64)
        else:
            if not isinstance(pts, torch.Tensor):
                raise RuntimeError("pts must be a tensor [sum n, 2] of (x, y)")
            if pts.dtype in (torch.int32, torch.int64, torch.int16):
                mode, pts = 1, pts.to(device=self.device, dtype=torch.int32).contiguous()
            else:
                mode, pts = 2, pts.to(device=self.device, dtype=torch.float32).contiguous()
            off = np.ascontiguousarray(offsets if offsets is not None else [0, pts.shape[0]], dtype=np.int64)
            if off.size != B + 1 or off[0] != 0 or off[-1] != pts.shape[0] or pts.dim() != 2 or pts.shape[1] != 2:
                raise RuntimeError("offsets must be [B+1] prefix sums covering pts [sum n, 2]")
            counts = np.diff(off)
            pp = pts
        return flow, Kinv, B, H, W, mode, pp, off, counts

    def flow_to_points(self, flow, Kinv, margin=10, pts=None, offsets=None):
        """flow [B,2,H,W] f32, Kinv [B,3,3] -> (x1, x2 f64 [sum n, 2], offsets [B+1]).
        pts None: dense crop with `margin`; int tensor [sum n,2] (x,y): pixel gather; float tensor:
        bilinear sampling (cfg.SAMPLE_SP)."""
        flow, Kinv, B, H, W, mode, pp, off, counts = self._flow_args(flow, Kinv, margin, pts, offsets)
        total = int(counts.sum())
        with torch.cuda.device(self.device):
            x1 = torch.empty(total, 2, dtype=torch.float64, device=self.device)
            x2 = torch.empty(total, 2, dtype=torch.float64, device=self.device)
            rc = self.L.tv5_flow_to_points(
                self.ctx, self._stream(), flow.data_ptr(), B, H, W, Kinv.data_ptr(), mode, int(margin),
                pp.data_ptr() if pp is not None else None,
                off.ctypes.data_as(C.POINTER(C.c_int64)) if off is not None else None,
                x1.data_ptr(), x2.data_ptr())
        _check(self.ctx, rc, "tv5_flow_to_points")
        return x1, x2, np.concatenate([[0], np.cumsum(counts)])

    def pose_from_flow(self, flow, Kinv, iters, thr, margin=10, pts=None, offsets=None, sets=None,
                       with_cheirality=True):
        """The geometric part of SFMnet.pose_by_ransac in one submission: returns
        (P_mat f32 [B,3,4], E_mat f32 [B,3,3], PoseResult with the float64 E/P and the counters)."""
        flow, Kinv, B, H, W, mode, pp, off, counts = self._flow_args(flow, Kinv, margin, pts, offsets)
        with torch.cuda.device(self.device):
            E32 = torch.empty(B, 3, 3, dtype=torch.float32, device=self.device)
            P32 = torch.empty(B, 3, 4, dtype=torch.float32, device=self.device)
            E = torch.empty(B, 3, 3, dtype=torch.float64, device=self.device)
            P = torch.empty(B, 3, 4, dtype=torch.float64, device=self.device)
            stats = torch.empty(B, 8, dtype=torch.int32, device=self.device)
            sp = None
            if sets is not None:
                sets = self._sets(sets, iters, B)
                sp = sets.data_ptr()
            rc = self.L.tv5_pose_from_flow(
                self.ctx, self._stream(), flow.data_ptr(), B, H, W, Kinv.data_ptr(), mode, int(margin),
                pp.data_ptr() if pp is not None else None,
                off.ctypes.data_as(C.POINTER(C.c_int64)) if off is not None else None,
                sp, int(iters), float(thr), int(bool(with_cheirality)), E32.data_ptr(), P32.data_ptr(),
                stats.data_ptr(), E.data_ptr(), P.data_ptr())
        _check(self.ctx, rc, "tv5_pose_from_flow")
        return P32, E32, PoseResult(E, P, stats)

    # -- plane-sweep cost volume (consumer of P: models/PSNet.py:141-157) -----------------------
    def plane_sweep(self, ref_fea, tgt_fea, pose, intrinsics4, intrinsics_inv4, nlabel, mindepth=1.0,
                    by_depth=False, out=None):
        """ref_fea, tgt_fea [B,C,h,w] f32; pose [B,3,4]; intrinsics at feature resolution [B,3,3]
        -> cost [B,2C,nlabel,h,w] f32 (what PSNet.forward builds before its 3-D convolutions)."""
        for t, name in ((ref_fea, "ref_fea"), (tgt_fea, "tgt_fea")):
            if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32 or t.dim() != 4:
                raise RuntimeError(f"{name} must be a CUDA float tensor [B, C, h, w]")
        if ref_fea.shape != tgt_fea.shape:
            raise RuntimeError("ref_fea and tgt_fea must have the same shape")
        B, Cc, h, w = ref_fea.shape
        ref_fea, tgt_fea = ref_fea.contiguous(), tgt_fea.contiguous()
        f32 = dict(device=self.device, dtype=torch.float32)
        pose = pose.to(**f32).contiguous()
        K = intrinsics4.to(**f32).contiguous()
        Ki = intrinsics_inv4.to(**f32).contiguous()
        if pose.numel() != B * 12 or K.numel() != B * 9 or Ki.numel() != B * 9:
            raise RuntimeError("pose must be [B,3,4] and the intrinsics [B,3,3]")
        with torch.cuda.device(self.device):
            if out is None:
                out = torch.empty(B, 2 * Cc, int(nlabel), h, w, **f32)
            elif out.shape != (B, 2 * Cc, int(nlabel), h, w) or out.dtype != torch.float32 or not out.is_contiguous():
                raise RuntimeError("out must be a contiguous float tensor [B, 2C, nlabel, h, w]")
            rc = self.L.tv5_plane_sweep(self.ctx, self._stream(), ref_fea.data_ptr(), tgt_fea.data_ptr(),
                                        pose.data_ptr(), K.data_ptr(), Ki.data_ptr(), B, Cc, h, w, int(nlabel),
                                        float(mindepth), int(bool(by_depth)), out.data_ptr())
        _check(self.ctx, rc, "tv5_plane_sweep")
        return out

    # -- decomposition and refinement (polish_E.cu in the reference) ---------------------------
    def decompose_batch(self, E, want_angles=True, want_uv=True):
        """E: [B,3,3] (or [3,3]) float64 CUDA -> dict(angles [B,5], U [B,3,3], V [B,3,3])."""
        if not E.is_cuda or E.dtype != torch.float64:
            raise RuntimeError("Emat must be a CUDA double tensor")
        Ef = E.contiguous().view(-1, 9)
        B = Ef.shape[0]
        with torch.cuda.device(self.device):
            ang = torch.empty(B, 5, dtype=torch.float64, device=self.device) if want_angles else None
            U = torch.empty(B, 3, 3, dtype=torch.float64, device=self.device) if want_uv else None
            V = torch.empty(B, 3, 3, dtype=torch.float64, device=self.device) if want_uv else None
            rc = self.L.tv5_decompose_batch(self.ctx, self._stream(), Ef.data_ptr(), B,
                                            ang.data_ptr() if want_angles else None,
                                            U.data_ptr() if want_uv else None,
                                            V.data_ptr() if want_uv else None)
        _check(self.ctx, rc, "tv5_decompose_batch")
        return dict(angles=ang, U=U, V=V)

    def optimise(self, x1, x2, E, delta, alpha, max_reps, mask=None, want_iters=False):
        """Refine E [3,3] on device points x1, x2 [N,2] (optionally only where mask != 0)."""
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        N = x1.shape[0]
        if x2.shape[0] != N:
            raise RuntimeError("input1 and input2 must have the same number of points")
        with torch.cuda.device(self.device):
            Eo = E.to(device=self.device, dtype=torch.float64).contiguous().clone().view(3, 3)
            it = torch.zeros(1, dtype=torch.int32, device=self.device) if want_iters else None
            if mask is not None:
                mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
                if mask.numel() != N:
                    raise RuntimeError("mask must have one entry per point")
            rc = self.L.tv5_optimise(self.ctx, self._stream(), x1.data_ptr(), x2.data_ptr(), N,
                                     mask.data_ptr() if mask is not None else None, Eo.data_ptr(),
                                     float(delta), float(alpha), int(max_reps),
                                     it.data_ptr() if want_iters else None)
        _check(self.ctx, rc, "tv5_optimise")
        return (Eo, it) if want_iters else Eo

    def optimise_batch(self, x1, x2, offsets, E, delta, alpha, max_reps, mask=None):
        """B problems: points concatenated, offsets [B+1] host prefix sums, E [B,3,3]."""
        _require_points(x1, "input1")
        _require_points(x2, "input2")
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        B = off.size - 1
        if B < 1 or off[0] != 0 or off[-1] != x1.shape[0] or x2.shape[0] != x1.shape[0]:
            raise RuntimeError("offsets must be prefix sums covering all points")
        with torch.cuda.device(self.device):
            Eo = E.to(device=self.device, dtype=torch.float64).contiguous().clone().view(B, 3, 3)
            it = torch.zeros(B, dtype=torch.int32, device=self.device)
            if mask is not None:
                mask = mask.to(device=self.device, dtype=torch.uint8).contiguous()
            rc = self.L.tv5_optimise_batch(self.ctx, self._stream(), B, x1.data_ptr(), x2.data_ptr(),
                                           off.ctypes.data_as(C.POINTER(C.c_int64)),
                                           mask.data_ptr() if mask is not None else None,
                                           Eo.data_ptr(), float(delta), float(alpha), int(max_reps),
                                           it.data_ptr())
        _check(self.ctx, rc, "tv5_optimise_batch")
        return Eo, it

    def optimise_host(self, x1, x2, E, delta, alpha, max_reps):
        """Host numpy in/out (the reference's calling convention: CPU tensors)."""
        x1 = np.ascontiguousarray(x1, dtype=np.float64)
        x2 = np.ascontiguousarray(x2, dtype=np.float64)
        E = np.ascontiguousarray(E, dtype=np.float64).reshape(9)
        out = np.empty(9)
        with torch.cuda.device(self.device):
            rc = self.L.tv5_optimise_host(self.ctx, self._stream(), x1.ctypes.data, x2.ctypes.data,
                                          int(x1.shape[0]), E.ctypes.data, float(delta), float(alpha),
                                          int(max_reps), out.ctypes.data)
        _check(self.ctx, rc, "tv5_optimise_host")
        return out.reshape(3, 3)

    # -- hypothesis-sharded single pair: device-side winner record / pick (tv5/dist.py) ---------
    RECORD_DOUBLES = 24   # TV5_WINNER_RECORD_BYTES / 8

    def winner_record(self, r, set_offset):
        """192-byte record of a single-pair PoseResult whose set ids start at `set_offset` in the
        global hypothesis table; float64 [24] on the device (bit container, see include/tv5.h)."""
        with torch.cuda.device(self.device):
            rec = torch.empty(self.RECORD_DOUBLES, dtype=torch.float64, device=self.device)
            rc = self.L.tv5_winner_record(self.ctx, self._stream(), r.E.data_ptr(), r.P.data_ptr(),
                                          r.stats.data_ptr(), int(set_offset), rec.data_ptr())
        _check(self.ctx, rc, "tv5_winner_record")
        return rec

    def winner_pick(self, records):
        """records: float64 [G*24] (all-gathered) -> PoseResult of the global first maximum."""
        G = records.numel() // self.RECORD_DOUBLES
        with torch.cuda.device(self.device):
            E = torch.empty(3, 3, dtype=torch.float64, device=self.device)
            P = torch.empty(3, 4, dtype=torch.float64, device=self.device)
            stats = torch.empty(8, dtype=torch.int32, device=self.device)
            rc = self.L.tv5_winner_pick(self.ctx, self._stream(), records.data_ptr(), G, E.data_ptr(),
                                        P.data_ptr(), stats.data_ptr())
        _check(self.ctx, rc, "tv5_winner_pick")
        return PoseResult(E, P, stats)

    # -- testing aid: guard zones + poisoned workspace (include/tv5.h: tv5_debug_guard) ---------
    def debug_guard(self, poison=0xFF):
        """Fresh engine only.  Every workspace buffer from now on: guard zones, payload = poison."""
        _check(self.ctx, self.L.tv5_debug_guard(self.ctx, 1, int(poison)), "tv5_debug_guard")

    def debug_poison(self, poison):
        _check(self.ctx, self.L.tv5_debug_poison(self.ctx, int(poison)), "tv5_debug_poison")

    def debug_stray_write(self, back=True):
        _check(self.ctx, self.L.tv5_debug_stray_write(self.ctx, int(bool(back))), "tv5_debug_stray_write")

    def debug_check_guards(self):
        """(guard bytes overwritten, guarded buffers); synchronises."""
        bad, n = C.c_int64(), C.c_int32()
        _check(self.ctx, self.L.tv5_debug_check_guards(self.ctx, C.byref(bad), C.byref(n)), "tv5_debug_check_guards")
        return bad.value, n.value

    # -- measurement --------------------------------------------------------------------------
    def measure_fp32_peak(self, mode=1):
        v = C.c_double()
        with torch.cuda.device(self.device):
            _check(self.ctx, self.L.tv5_measure_fp32_peak(self.ctx, int(mode), C.byref(v)), "fp32 peak")
        return v.value

    def set_force_exact(self, on=True):
        _check(self.ctx, self.L.tv5_set_force_exact(self.ctx, int(bool(on))), "set_force_exact")

    def set_graphs(self, on=True):
        """CUDA-graph replay of single-pair submissions (default on)."""
        _check(self.ctx, self.L.tv5_set_graphs(self.ctx, int(bool(on))), "set_graphs")

    def set_early_exit(self, on=True):
        """Staged scoring with exact hypothesis pruning (default off); results do not depend on it."""
        _check(self.ctx, self.L.tv5_set_early_exit(self.ctx, int(bool(on))), "set_early_exit")
        self.early_exit = bool(on)

    def set_split_solver(self, on=True):
        """Three-kernel solver (default) vs the fused kernel; results do not depend on it."""
        _check(self.ctx, self.L.tv5_set_split_solver(self.ctx, int(bool(on))), "set_split_solver")

    def set_overlap(self, on=True):
        """Solver/scorer overlap inside one submission (default off); results do not depend on it."""
        _check(self.ctx, self.L.tv5_set_overlap(self.ctx, int(bool(on))), "set_overlap")
        self.overlap = bool(on)

    def pipeline_chunks(self, B):
        """Chunks a submission of B pairs is cut into (tv5_internal.h: kPipeChunks, kPipeMinPairs)."""
        return max(1, min(8, int(B) // 16)) if getattr(self, "overlap", False) else 1

    def profile_enable(self, on=True):
        _check(self.ctx, self.L.tv5_profile_enable(self.ctx, int(bool(on))), "profile_enable")

    def profile_read(self, reset=True):
        ms = (C.c_double * _lib.TV5_N_STAGES)()
        n = (C.c_int64 * _lib.TV5_N_STAGES)()
        _check(self.ctx, self.L.tv5_profile_read(self.ctx, ms, n, int(bool(reset))), "profile_read")
        return {name: (ms[i], n[i]) for i, name in enumerate(_lib.STAGE_NAMES)}


_engines = {}


def get_engine(device=None):
    """Cached Engine of a device (default: current CUDA device)."""
    if not torch.cuda.is_available():
        raise Tv5Error("tv5 needs a CUDA device (sm_100a); there is no CPU fallback")
    idx = torch.cuda.current_device() if device is None else (torch.device(device).index or 0)
    if idx not in _engines:
        _engines[idx] = Engine(torch.device("cuda", idx))
    return _engines[idx]


def compute_pose(x1, x2, iters, thr, **kw):
    return get_engine(x1.device).compute_pose(x1, x2, iters, thr, **kw)


def compute_pose_batch(x1, x2, offsets, iters, thr, **kw):
    return get_engine(x1.device).compute_pose_batch(x1, x2, offsets, iters, thr, **kw)


def solve5(x1, x2, sets, **kw):
    return get_engine(x1.device).solve5(x1, x2, sets, **kw)


def score(x1, x2, E_list, thr, **kw):
    return get_engine(x1.device).score(x1, x2, E_list, thr, **kw)


def score_bounds(x1, x2, E_list, thr, **kw):
    return get_engine(x1.device).score_bounds(x1, x2, E_list, thr, **kw)


def decompose_host(E):
    """Five Givens angles of E (host 3x3) — tv5_decompose; needs libtv5.so but no GPU."""
    E = np.ascontiguousarray(E, dtype=np.float64).reshape(9)
    out = np.empty(5)
    _check(None, _lib.load_library().tv5_decompose(E.ctypes.data, out.ctypes.data), "tv5_decompose")
    return out


def decompose_uv_host(E):
    E = np.ascontiguousarray(E, dtype=np.float64).reshape(9)
    U, V = np.empty(9), np.empty(9)
    _check(None, _lib.load_library().tv5_decompose_uv(E.ctypes.data, U.ctypes.data, V.ctypes.data),
           "tv5_decompose_uv")
    return U.reshape(3, 3), V.reshape(3, 3)


def ref_rng_sets(N, iters, device=None):
    return get_engine(device).ref_rng_sets(N, iters)
