"""The UNMODIFIED reference (kamwoh/DeepIPR), made available where /root/reference does not exist — TEST / BASELINE
INFRASTRUCTURE ONLY (same rule as the rest of oracle/: tests/, __graft_entry__.smoke() and bench.py's baseline legs).

The reference is a pure-Python script tree (no setup.py, nothing to compile), so "building" it means packing its
sources into ONE archive, `oracle/_ref/deepipr_reference.zip`, from where they lie under /root/reference.  The archive
is a build output like the compiled .so files: git-ignored (never part of the history — no reference source is copied
into the repository), not gpurun-ignored (so it travels to the GPU box, which has no /root/reference).  At run time
the archive is unpacked into a temporary directory and imported from there.

  build()              pack /root/reference -> oracle/_ref/deepipr_reference.zip   (__graft_entry__.build())
  locate()             directory holding an importable reference tree (checkout, else unpacked archive), or None
  workdir()            a fresh writable copy (the reference's scripts write logs/ into the cwd)
  import_reference()   import the reference's model / trainer modules, stock or with this repo's blocks patched in
"""
import importlib
import os
import shutil
import sys
import tempfile
import zipfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.environ.get("DEEPIPR_REFERENCE", "/root/reference")
BUNDLE_DIR = os.path.join(HERE, "_ref")
BUNDLE = os.path.join(BUNDLE_DIR, "deepipr_reference.zip")
_KEEP_EXT = (".py", ".json", ".sh", ".txt", ".yml", ".md")
_extracted = None


def _has_tree(path):
    return bool(path) and os.path.isdir(os.path.join(path, "models")) and os.path.isdir(os.path.join(path, "experiments"))


def build(force=False):
    """Pack the reference checkout into the archive.  No-op (keeps an existing archive) when there is no checkout."""
    if not _has_tree(SRC):
        return BUNDLE if os.path.exists(BUNDLE) else None
    files = []
    for base, dirs, names in os.walk(SRC):
        dirs[:] = sorted(d for d in dirs if d not in (".git", "__pycache__", "docs", "logs", "data"))
        for n in sorted(names):
            if n.endswith(_KEEP_EXT) or n == "Dockerfile":
                files.append(os.path.join(base, n))
    os.makedirs(BUNDLE_DIR, exist_ok=True)
    newest = max(os.path.getmtime(f) for f in files)
    if not force and os.path.exists(BUNDLE) and os.path.getmtime(BUNDLE) >= newest:
        with zipfile.ZipFile(BUNDLE) as z:
            if len(z.namelist()) == len(files):
                return BUNDLE
    tmp = BUNDLE + ".tmp"
    with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
        for f in files:
            info = zipfile.ZipInfo(os.path.relpath(f, SRC), date_time=(2020, 1, 1, 0, 0, 0))   # reproducible archive
            info.compress_type = zipfile.ZIP_DEFLATED
            info.external_attr = 0o644 << 16
            with open(f, "rb") as fh:
                z.writestr(info, fh.read())
    os.replace(tmp, BUNDLE)
    return BUNDLE


def _extract():
    global _extracted
    if _extracted and _has_tree(_extracted):
        return _extracted
    shared = os.environ.get("DEEPIPR_REFERENCE_EXTRACTED")
    if _has_tree(shared):
        _extracted = shared
        return _extracted
    if not os.path.exists(BUNDLE):
        return None
    d = tempfile.mkdtemp(prefix="deepipr_ref_")
    with zipfile.ZipFile(BUNDLE) as z:
        z.extractall(d)
    _extracted = d
    os.environ["DEEPIPR_REFERENCE_EXTRACTED"] = d      # child processes reuse it
    return d


def locate(prefer_bundle=False):
    """Read-only directory to import the reference from.  prefer_bundle=True exercises the route the GPU box takes."""
    if prefer_bundle or not _has_tree(SRC):
        d = _extract()
        if d:
            return d
    return SRC if _has_tree(SRC) else None


def available():
    return locate() is not None


def workdir(prefer_bundle=True):
    """Fresh writable copy of the tree (caller deletes it)."""
    src = locate(prefer_bundle)
    if src is None:
        raise RuntimeError("the reference is neither checked out nor bundled (run __graft_entry__.build() where "
                           "/root/reference exists)")
    d = tempfile.mkdtemp(prefix="deepipr_refwork_")
    shutil.copytree(src, d, dirs_exist_ok=True, ignore=shutil.ignore_patterns(".git", "__pycache__", "logs", "data"))
    return d


_REF_TOPLEVEL = ("models", "experiments", "dataset", "passport_generator")


def _purge_modules():
    saved = {}
    for name in list(sys.modules):
        if name.split(".")[0] in _REF_TOPLEVEL:
            saved[name] = sys.modules.pop(name)
    return saved


def import_reference(patched, names=("models.resnet_passport_private", "models.resnet_passport", "models.resnet_normal",
                                     "models.alexnet_passport", "models.alexnet_normal",
                                     "experiments.trainer_private", "experiments.trainer", "experiments.utils"),
                     path=None):
    """Import reference modules and return {module name: module}.  patched=False: the stock reference (its own
    PassportBlock / ConvBlock: torch eager).  patched=True: after deepipr_b200.patch_reference(), i.e. the
    reference's model and trainer files running on this repository's blocks.  Both flavours can live in one process:
    the module table is swapped around the import and restored afterwards."""
    path = path or locate()
    if path is None:
        raise RuntimeError("reference not available")
    outer = _purge_modules()
    sys.path.insert(0, path)
    try:
        if patched:
            if ROOT not in sys.path:
                sys.path.insert(0, ROOT)
            import deepipr_b200
            deepipr_b200.patch_reference()
        mods = {n: importlib.import_module(n) for n in names}
    finally:
        sys.path.remove(path)
        _purge_modules()
        sys.modules.update(outer)
    return mods
