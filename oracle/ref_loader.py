"""TEST / BENCH INFRASTRUCTURE ONLY -- never imported by the product (pixel-embedded-affinity_b200/).

Runs the UNMODIFIED reference (weih527/Pixel-Embedded-Affinity) for the hot path.  The reference is flat Python with
cwd-relative imports and cannot be pip-installed; `/root/reference` does not exist on the GPU box.  So:

  * ``stage()`` (called by ``__graft_entry__.build()`` in the build container, where /root/reference exists) copies the
    handful of reference files the path needs -- loss functions, criteria, GT generation, the two backbones of BASELINE
    config 5 and the consistency flip -- verbatim into the git-ignored ``baseline/_ref/`` tree.  That tree travels to
    the GPU box with the gpurun snapshot like the built .so does; it is never committed.
  * ``load()`` imports them by path from ``baseline/_ref`` (or /root/reference when present) with the two shims
    SURVEY.md 8(c) documents: a no-op ``Tensor.cuda`` when the process has no CUDA device (WeightedMSE calls ``.cuda()``
    on its normaliser, scripts_cvppp/loss/loss.py:116) and stub ``skimage`` modules for data_segmentation.py:4-6.

``bench.py --impl reference`` and the ``cpu_baseline`` leg use it when the staged tree exists (``kind: "reference"``)
and fall back to the oracle port (oracle/pea_oracle.py, pinned to the reference bit for bit by tests/golden) otherwise.
"""
import importlib.util
import os
import shutil
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref")
SOURCE = os.environ.get("PEA_REFERENCE", "/root/reference")

FILES = [
    "scripts_cvppp/loss/loss_embedding_mse.py",
    "scripts_cvppp/loss/loss.py",
    "scripts_cvppp/utils/affinity_ours.py",
    "scripts_cvppp/data/data_affinity.py",
    "scripts_cvppp/data/data_segmentation.py",
    "scripts_cvppp/data/data_transform.py",   # imported by data_segmentation.py:9
    "scripts_cvppp/data/data_misc.py",        # imported by data_transform.py:11
    "scripts_cvppp/data/data_consistency.py",
    "scripts_cvppp/model/unet2d_residual.py",
    "scripts_ac3ac4/loss/loss_embedding_mse.py",
    "scripts_ac3ac4/loss/loss.py",
    "scripts_ac3ac4/utils/consistency_aug.py",
    "scripts_ac3ac4/model/model_superhuman.py",
    "scripts_ac3ac4/model/basic.py",
    "scripts_ac3ac4/model/residual.py",
    "scripts_ac3ac4/model/model_para.py",
]


def stage(source=SOURCE, dest=STAGED):
    """Copy the reference files of the path into baseline/_ref (no-op when the reference tree is absent)."""
    if not os.path.isdir(source):
        return False
    for rel in FILES:
        src, dst = os.path.join(source, rel), os.path.join(dest, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
    return True


def root():
    """Directory holding the reference files: the staged copy, else the reference tree itself, else None."""
    for cand in (STAGED, SOURCE):
        if os.path.exists(os.path.join(cand, FILES[0])):
            return cand
    return None


def available():
    return root() is not None


def _load(name, base, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(base, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(models=False):
    """Import the reference modules by path.  Returns a namespace with loss2d, loss3d, crit, aff2d, aff3d, seg,
    flip2d, flip3d (and unet2d / unet3d with ``models=True``)."""
    import torch
    base = root()
    if base is None:
        raise FileNotFoundError("reference files not staged: run __graft_entry__.build() where /root/reference exists")
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self   # GPU-less host; on a GPU box callers scope the shim themselves
    for n in ("skimage", "skimage.morphology", "skimage.measure", "skimage.segmentation"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__getattr__ = lambda attr: None
            sys.modules[n] = m
    ref = types.SimpleNamespace(root=base)
    generic = ("data", "model", "loss", "utils")  # the reference's cwd-relative package names

    def scoped(tree, loads):
        """Import with `<base>/<tree>` on sys.path (the reference's scripts run from their own directory), then take the
        generic package names back out of sys.modules so that the two script trees cannot see each other's packages."""
        saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split(".")[0] in generic}
        sys.path.insert(0, os.path.join(base, tree))
        try:
            for attr, name, rel in loads:
                setattr(ref, attr, _load(name, base, rel))
        finally:
            sys.path.pop(0)
            for k in [k for k in sys.modules if k.split(".")[0] in generic]:
                del sys.modules[k]
            sys.modules.update(saved)

    scoped("scripts_cvppp", [
        ("loss2d", "pea_ref_loss2d", "scripts_cvppp/loss/loss_embedding_mse.py"),
        ("crit", "pea_ref_crit", "scripts_cvppp/loss/loss.py"),
        ("aff2d", "pea_ref_aff2d", "scripts_cvppp/utils/affinity_ours.py"),
        ("aff3d", "pea_ref_aff3d", "scripts_cvppp/data/data_affinity.py"),
        ("seg", "pea_ref_seg", "scripts_cvppp/data/data_segmentation.py"),
        ("flip2d", "pea_ref_flip2d", "scripts_cvppp/data/data_consistency.py"),
    ])
    loads3 = [("loss3d", "pea_ref_loss3d", "scripts_ac3ac4/loss/loss_embedding_mse.py")]
    try:
        import cv2  # noqa: F401  (consistency_aug.py imports it; only the 3-D flip needs that module)
        loads3.append(("flip3d", "pea_ref_flip3d", "scripts_ac3ac4/utils/consistency_aug.py"))
    except ImportError:
        ref.flip3d = None
    scoped("scripts_ac3ac4", loads3)
    if models:
        scoped("scripts_cvppp", [("unet2d", "pea_ref_unet2d", "scripts_cvppp/model/unet2d_residual.py")])
        scoped("scripts_ac3ac4", [("unet3d", "pea_ref_unet3d", "scripts_ac3ac4/model/model_superhuman.py")])
    return ref
