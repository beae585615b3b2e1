"""configs[4] on the GPU box: the reference's SFMnet.forward (staged copy, unmodified) with the
pose stage served by this repo's drop-in module vs by the compiled reference extension.
    python tools/config5_check.py [--nlabel 128] [--reps 3] [--out gpurun_out/config5.json]
"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "deep-sfm-revisited_b200"), os.path.join(ROOT, "baseline")):
    if p not in sys.path:
        sys.path.insert(0, p)
import numpy as np
import torch
import harness, scene
from tv5 import synth


class FixedFlow(torch.nn.Module):
    """Stands in for the flow network when a meaningful flow is wanted (random-init DICL outputs noise)."""
    def __init__(self, flow, pad_hw):
        super().__init__()
        f = torch.zeros(1, 2, *pad_hw)
        f[0, :, :flow.shape[1], :flow.shape[2]] = torch.from_numpy(flow)
        self.register_buffer("flow", f)
    def forward(self, x):
        return self.flow.clone(), torch.ones_like(self.flow[:, :1])


def run_forward(net, sc, timings=None):
    H, W = sc["ref"].shape[1:]
    Hp, Wp = int(np.ceil(H / 128) * 128), int(np.ceil(W / 128) * 128)
    ref = torch.from_numpy(sc["ref"])[None].cuda()
    tgt = torch.from_numpy(sc["target"])[None].cuda()
    pad = (0, Wp - W, 0, Hp - H)
    ref = torch.nn.functional.pad(ref, pad, "replicate")
    tgt = torch.nn.functional.pad(tgt, pad, "replicate")
    K = torch.from_numpy(sc["K"])[None]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.no_grad():
        flow, P, depth, _ = net(ref, tgt, K, None, None, False, H, W)
    torch.cuda.synchronize()
    return flow, P, depth, (time.perf_counter() - t0) * 1e3


def stage_timers(net):
    """wall-clock (with device sync) of the three stages of SFMnet.forward, via instance wrappers"""
    rec = {"flow_ms": [], "pose_ms": [], "depth_ms": []}
    def wrap(fn, key):
        def w(*a, **k):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            r = fn(*a, **k)
            torch.cuda.synchronize(); rec[key].append((time.perf_counter() - t0) * 1e3)
            return r
        return w
    net.flow_estimator.forward = wrap(net.flow_estimator.forward, "flow_ms")
    net.pose_by_ransac = wrap(net.pose_by_ransac, "pose_ms")
    net.depth_estimator.forward = wrap(net.depth_estimator.forward, "depth_ms")
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nlabel", type=int, default=128)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "config5.json"))
    a = ap.parse_args()
    out = {}
    ref = harness.load_reference("tv5")
    sc = scene.make_scene(0)
    H, W = sc["ref"].shape[1:]
    Hp, Wp = int(np.ceil(H / 128) * 128), int(np.ceil(W / 128) * 128)
    have_ref = harness.refext_path() is not None
    for variant in ("synthetic_flow", "random_init_dicl"):
        net = ref.make_sfmnet(a.nlabel, seed=0)
        if variant == "synthetic_flow":
            net.flow_estimator = FixedFlow(sc["flow"], (Hp, Wp)).cuda()
        rec = stage_timers(net)
        res = {}
        for be in (("tv5", "refext") if have_ref else ("tv5",)):
            ref.use_backend(be)
            for k in rec: rec[k].clear()
            runs = []
            for r in range(a.reps):
                flow, P, depth, ms = run_forward(net, sc)
                runs.append(ms)
            Pn = P[0, 0].double().cpu().numpy()
            res[be] = {"P": Pn.tolist(), "total_ms": runs, "stage_ms": {k: list(v) for k, v in rec.items()},
                       "rot_err_deg": synth.rotation_error_deg(Pn[:, :3], sc["R"]),
                       "trans_err_deg": synth.translation_error_deg(Pn[:, 3] / (np.linalg.norm(Pn[:, 3]) + 1e-30), sc["t"]),
                       "depth_finite": bool(torch.isfinite(depth).all()), "depth_mean": float(depth.mean())}
            res[be]["_depth"] = depth
        if have_ref:
            d0, d1 = res["tv5"].pop("_depth"), res["refext"].pop("_depth")
            P0, P1 = np.array(res["tv5"]["P"]), np.array(res["refext"]["P"])
            res["P_max_abs_diff"] = float(np.abs(P0 - P1).max())
            res["depth_max_abs_diff"] = float((d0 - d1).abs().max())
            res["depth_rel_diff"] = float(((d0 - d1).abs() / d1.abs().clamp_min(1e-6)).max())
        else:
            res["tv5"].pop("_depth")
        out[variant] = res
        print(variant, json.dumps({k: v for k, v in res.items()}, default=str)[:1500], flush=True)
        del net
        torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(out, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
