"""Import shims that let the reference's own Python code for the hot path run in THIS container
(CPU only, Python 3.12, none of gsplat / nvdiffrast / tinycudann / open3d installed).

Used only by scripts/make_golden.py to generate tests/golden/*.npz.  /root/reference is read-only and
is not present on the GPU box, so nothing under tests/ or bench.py imports this at run time.
Technique from SURVEY.md section 8(c): pre-register `sys.modules` stand-ins for the missing
third-party names and for the two import-time-JIT native plugins, then exec the first 481 lines of
rfstudio/model/geosplat.py (RenderableAttrs, MGAdapter, tone mapping) from source.
"""
import importlib.machinery
import sys
import types

REF = "/root/reference"


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __mro_entries__(self, bases):
        return (object,)

    def __getitem__(self, item):
        return _Anything()

    def __or__(self, other):
        return _Anything()

    __ror__ = __or__


def _stub(name, attrs=None, package=True):
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=package)
    if package:
        m.__path__ = []

    def _getattr(attr):
        if attr.startswith("__"):
            raise AttributeError(attr)
        return _Anything()

    m.__getattr__ = _getattr
    for k, v in (attrs or {}).items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install(texture_fn, rasterization_fn=None):
    """texture_fn(tex, uv, mip=None, mip_level_bias=None, filter_mode=..., boundary_mode=...) stands in for
    nvdiffrast.torch.texture; rasterization_fn for gsplat.rasterization."""
    import numpy  # noqa: F401  (real modules first: stubs must never shadow them)
    import torch  # noqa: F401
    import torch.nn.functional  # noqa: F401
    if REF not in sys.path:
        sys.path.insert(0, REF)
    for name in ["open3d", "open3d.t", "open3d.t.geometry", "open3d.t.io", "kornia", "kornia.filters", "trimesh",
                 "pyexr", "rfviser", "rfviser.transforms", "plotext", "skimage", "skimage.measure", "matplotlib",
                 "matplotlib.pyplot", "matplotlib.cm", "torchmetrics", "torchmetrics.image", "torchmetrics.functional",
                 "torchmetrics.image.lpip", "nerfacc", "tinycudann", "cv2", "sklearn", "sklearn.neighbors", "lpips",
                 "ffmpegcv", "imageio", "PIL", "PIL.Image", "mediapy", "viser", "tyro", "tyro.conf", "tyro.extras",
                 "pytorch3d", "pytorch3d.ops", "xatlas", "pymeshlab", "scipy.spatial", "gdown"]:
        try:
            __import__(name)
        except Exception:
            _stub(name)
    _stub("nvdiffrast")
    _stub("nvdiffrast.torch", {"texture": texture_fn})
    sys.modules["nvdiffrast"].torch = sys.modules["nvdiffrast.torch"]
    _stub("gsplat", {"rasterization": rasterization_fn or _Anything(), "rasterization_2dgs": _Anything()})
    _stub("rfstudio.graphics._mesh._optix", {"OptiXContext": _Anything, "bilateral_denoiser": _Anything(),
                                             "optix_env_shade": _Anything()})
    _stub("rfstudio.graphics._mesh._splitsum", {"diffuse_cubemap": _Anything(), "specular_cubemap": _Anything()})


def load_geosplat_head():
    """exec rfstudio/model/geosplat.py lines 1..481 (everything before the Python-3.12-incompatible
    dataclass at :482) with its two relative imports dropped.  Returns the namespace."""
    src = open(f"{REF}/rfstudio/model/geosplat.py").read().split("\n")[:481]
    src = [l for l in src if not l.startswith("from .")]
    mod = types.ModuleType("rfstudio_geosplat_head")
    sys.modules[mod.__name__] = mod
    ns = mod.__dict__
    ns["HashEncoding"] = _Anything
    ns["GSplatter"] = _Anything
    exec(compile("\n".join(src), f"{REF}/rfstudio/model/geosplat.py", "exec"), ns)
    return ns
