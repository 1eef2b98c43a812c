"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference under /root/reference.

Only usable inside the build container (the GPU box has no /root/reference).  Used by
oracle/make_golden.py to generate tests/golden/* and by tests marked `needs_reference`
to validate the restatements in oracle/.  Nothing in the product imports this.

Three harness-side shims, none of which edits the reference:
  1. combine_detect.py:9,17 import `ultralytics` and `foreign.recordDeal` at module top
     -> stub modules are pre-seeded in sys.modules.
  2. detect_face/utils/utils_bbox.py:19 calls `.cuda()` unconditionally
     -> `torch.Tensor.cuda` is neutralised to the identity *inside this process* when no GPU
        is present, so the reference's own function body runs unmodified on CPU.
  3. No weights ship -> callers pass a generated state_dict path (oracle/weights_oracle.py).
"""
import os
import sys
import types
import importlib

REFERENCE_ROOT = os.environ.get("VD_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "detect_face", "face.py"))


_loaded = {}


def load_reference():
    """Returns (ref_face_module, ref_combine_detect_module) imported from /root/reference."""
    if _loaded:
        return _loaded["face"], _loaded["combine"]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    import torch

    # shim 1: stub the imports that cannot load here
    ul = types.ModuleType("ultralytics")
    ul.YOLO = object
    sys.modules.setdefault("ultralytics", ul)
    foreign = types.ModuleType("foreign")
    rd = types.ModuleType("foreign.recordDeal")
    foreign.recordDeal = rd
    # the reference tree has a real `foreign/` dir with .so blobs for CPython 3.10; never load them
    saved = {k: sys.modules.get(k) for k in ("foreign", "foreign.recordDeal", "detect_face",
                                              "combine_detect")}
    sys.modules["foreign"] = foreign
    sys.modules["foreign.recordDeal"] = rd

    # shim 2: CPU-only box -> make the hard-coded .cuda() a no-op
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self

    # import the reference under private names so it cannot collide with the product's
    # drop-in `detect_face` / `combine_detect` modules that tests also import
    for k in [k for k in sys.modules if k == "detect_face" or k.startswith("detect_face.")
              or k == "combine_detect"]:
        saved[k] = sys.modules.pop(k)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        face = importlib.import_module("detect_face.face")
        combine = importlib.import_module("combine_detect")
        ref_mods = {k: v for k, v in sys.modules.items()
                    if k == "detect_face" or k.startswith("detect_face.") or k == "combine_detect"}
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for k in list(sys.modules):
            if k == "detect_face" or k.startswith("detect_face.") or k == "combine_detect":
                del sys.modules[k]
        for k, v in saved.items():
            if v is not None and k not in ("foreign", "foreign.recordDeal"):
                sys.modules[k] = v
    _loaded["face"], _loaded["combine"], _loaded["mods"] = face, combine, ref_mods
    return face, combine
