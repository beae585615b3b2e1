"""TEST / BENCH INFRASTRUCTURE.  Runs the reference's own Python — `epipolar_utils.py` and
`models/SFMnet.py`, byte for byte as staged by baseline/stage_ref_py.py — with the pose stage
behind `import essential_matrix` served either by this repo's drop-in module or by the compiled
reference extension (oracle/_ref/refext).  Nothing of the reference is edited: the modules it
cannot import in this image are stubbed in `sys.modules` and its global `cfg` is
filled from its own cfgs/kitti.yml through its own merge function.

    ref = load_reference("tv5")           # or "refext"
    ref.epipolar_utils.compute_P_matrix_ransac(...)
    net = ref.make_sfmnet(nlabel=128)     # random-init DICL + PSNet (no checkpoints offline)
    ref.use_backend("refext")             # swap what `essential_matrix.computeP` resolves to

Stubs injected (each only when the real module is missing):
  easydict.EasyDict      attribute dict, recursive (lib/config.py:1)
  path.Path              pathlib-backed (utils.py:5; only used by the trainer's folder naming)
  cv2.xfeatures2d        SIFT_create -> cv2.SIFT_create; SURF_create -> SIFT with a lower contrast
                         threshold (SURF is a non-free module absent from cv2 4.13; it is only the
                         "too few keypoints" fallback, models/SFMnet.py:205-208)
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PY = os.path.join(HERE, "_ref", "py")
PKG = os.path.join(ROOT, "deep-sfm-revisited_b200")
REFEXT_DIR = os.path.join(ROOT, "oracle", "_ref", "refext")


def staged():
    return os.path.exists(os.path.join(PY, ".staged"))


def refext_path():
    if not os.path.isdir(REFEXT_DIR):
        return None
    so = [f for f in os.listdir(REFEXT_DIR) if f.endswith(".so")]
    return os.path.join(REFEXT_DIR, so[0]) if so else None


# ---------------------------------------------------------------------------------------------
# stubs
# ---------------------------------------------------------------------------------------------
class _EasyDict(dict):
    """Minimal stand-in for easydict.EasyDict: keys are attributes, nested dicts are wrapped."""

    def __init__(self, d=None, **kw):
        super().__init__()
        for k, v in dict(d or {}, **kw).items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        elif isinstance(v, (list, tuple)):
            v = type(v)(_EasyDict(x) if isinstance(x, dict) and not isinstance(x, _EasyDict) else x for x in v)
        super().__setitem__(k, v)
        super().__setattr__(k, v)

    __setitem__ = __setattr__


def _install_stubs():
    try:
        import easydict  # noqa: F401
    except ImportError:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    try:
        import path  # noqa: F401
    except ImportError:
        import pathlib
        m = types.ModuleType("path")

        class Path(type(pathlib.Path())):
            def normpath(self):
                return Path(os.path.normpath(str(self)))

            def makedirs_p(self):
                os.makedirs(str(self), exist_ok=True)
                return self

        m.Path = Path
        sys.modules["path"] = m
    import cv2
    if not hasattr(cv2, "xfeatures2d"):
        ns = types.SimpleNamespace()
        ns.SIFT_create = lambda *a, **k: cv2.SIFT_create(*a, **k)
        ns.SURF_create = lambda *a, **k: cv2.SIFT_create(contrastThreshold=0.01)
        cv2.xfeatures2d = ns


# ---------------------------------------------------------------------------------------------
# backends of `import essential_matrix`
# ---------------------------------------------------------------------------------------------
_backends = {}


def backend(name):
    """'tv5': this repo's drop-in module; 'refext': the compiled reference extension."""
    if name in _backends:
        return _backends[name]
    if name == "tv5":
        for p in (PKG,):
            if p not in sys.path:
                sys.path.insert(0, p)
        saved = sys.modules.pop("essential_matrix", None)
        try:
            spec = importlib.util.spec_from_file_location(
                "essential_matrix", os.path.join(PKG, "essential_matrix", "__init__.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            if saved is not None:
                sys.modules["essential_matrix"] = saved
    elif name == "refext":
        so = refext_path()
        if so is None:
            raise RuntimeError("oracle/_ref/refext is not built (oracle/build_ref.sh ext)")
        import torch  # noqa: F401  (the extension links against libtorch)
        spec = importlib.util.spec_from_file_location("essential_matrix", so)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    else:
        raise ValueError(name)
    _backends[name] = mod
    return mod


class Reference:
    """The imported reference modules + the knob that swaps its `essential_matrix`."""

    def __init__(self, epipolar_utils, sfmnet_mod, cfg):
        self.epipolar_utils = epipolar_utils
        self.sfmnet_mod = sfmnet_mod
        self.cfg = cfg
        self.backend_name = None

    def use_backend(self, name):
        mod = backend(name)
        # the reference resolves `essential_matrix.<fn>` through these two module globals
        # (epipolar_utils.py:4, models/SFMnet.py:11) at call time
        self.epipolar_utils.essential_matrix = mod
        if self.sfmnet_mod is not None:
            self.sfmnet_mod.essential_matrix = mod
        sys.modules["essential_matrix"] = mod
        self.backend_name = name
        return mod

    def make_sfmnet(self, nlabel=128, seed=0, device="cuda"):
        """SFMnet exactly as main.py:198 builds it, random-init weights (seeded), eval mode."""
        import torch
        torch.manual_seed(seed)
        net = self.sfmnet_mod.SFMnet(nlabel)
        return net.to(device).eval()


_loaded = None


def load_reference(backend_name="tv5", with_models=True, yaml_name="kitti.yml", overrides=None):
    """Import the staged reference Python with `essential_matrix` = backend_name."""
    global _loaded
    if not staged():
        raise RuntimeError("reference Python tree not staged: run baseline/stage_ref_py.py where "
                           "/root/reference exists (__graft_entry__.build() does)")
    if _loaded is None:
        _install_stubs()
        if PY not in sys.path:
            sys.path.insert(0, PY)
        sys.modules["essential_matrix"] = backend(backend_name)
        cfgmod = importlib.import_module("lib.config")
        cfg = cfgmod.cfg
        if yaml_name:
            # what cfg_from_file (lib/config.py:380-386) does, with a Loader (PyYAML >= 6 needs one)
            import yaml
            with open(os.path.join(PY, "cfgs", yaml_name), encoding="utf-8-sig") as fh:
                ycfg = yaml.safe_load(fh)
            edict = sys.modules["easydict"].EasyDict
            # GT_DEPTH_DIR is a dataset path (None by default: the reference's own merge rejects the
            # placeholder string of its yaml); unknown keys would raise KeyError there
            known = {k: v for k, v in ycfg.items() if k in cfg and k != "GT_DEPTH_DIR"}
            cfgmod._merge_a_into_b(edict(known), cfg)
        eu = importlib.import_module("epipolar_utils")
        sm = importlib.import_module("models.SFMnet") if with_models else None
        _loaded = Reference(eu, sm, cfg)
    elif with_models and _loaded.sfmnet_mod is None:
        _loaded.sfmnet_mod = importlib.import_module("models.SFMnet")
    for k, v in (overrides or {}).items():
        setattr(_loaded.cfg, k, v)
    _loaded.use_backend(backend_name)
    return _loaded
