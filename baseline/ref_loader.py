"""Imports the UNMODIFIED reference (installed by baseline/build_ref.sh into baseline/_ref/) for baseline timing
and golden generation on the GPU box.  Third-party modules the reference imports at module scope but never uses on
the hot path (open3d, matplotlib, plyfile, e3nn, einsum, kornia, trimesh) are absent from this image and are
stubbed in sys.modules.  Nothing here is used by the product path."""
import importlib
import os
import sys
import types

REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def available() -> bool:
    d = os.path.join(REF_DIR, "diff_surfel_rasterization")
    return os.path.isdir(d) and any(f.startswith("_C") and f.endswith(".so") for f in os.listdir(d))


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(self.__name__ + "." + name)

    def __call__(self, *a, **k):
        return _Stub(self.__name__ + "()")


def _install_stubs():
    for name in ["open3d", "matplotlib", "matplotlib.pyplot", "matplotlib.colors", "plyfile", "e3nn", "e3nn.o3",
                 "einsum", "kornia", "trimesh", "mediapy", "lpips", "cv2"]:
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = _Stub(name)


def load():
    """Returns (render, contrastive_loss, diff_surfel_rasterization module) of the reference."""
    if not available():
        raise RuntimeError("baseline/_ref is missing: run baseline/build_ref.sh where /root/reference exists")
    _install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    dsr = importlib.import_module("diff_surfel_rasterization")
    gr = importlib.import_module("gaussian_renderer")
    cu = importlib.import_module("utils.contrastive_utils")
    return gr.render, cu.contrastive_loss, dsr
