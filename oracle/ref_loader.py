"""Loads the reference's OWN functions.  TEST INFRASTRUCTURE ONLY.

Source tree, in this order: /root/reference (the build container) or ``oracle/_ref/`` (byte-for-byte copies made
by ``oracle/build_ref.py``; git-ignored, shipped to the GPU box where /root/reference does not exist).

``datautils/utils.py`` imports cleanly (torch + heapq only).  ``extract_features.py`` does not (h5py / cv2 missing), so
``sample_frame_indices`` (``extract_features.py:32-39``) is compiled at run time from that file's own AST node --
nothing is copied into this repository's history.
"""
from __future__ import annotations

import ast
import importlib.util
import os

from . import build_ref

REFERENCE_ROOT = os.environ.get("SASVQA_REFERENCE_ROOT", "/root/reference")


def source_root() -> str | None:
    """Directory that holds the reference's ``src/`` tree, or None."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "preprocessing", "datautils", "utils.py")):
        return REFERENCE_ROOT
    if build_ref.available():
        return build_ref.REF_OUT
    return None


def source_kind() -> str:
    root = source_root()
    if root is None:
        return "absent"
    return "/root/reference" if root == REFERENCE_ROOT else "oracle/_ref"


def available() -> bool:
    return source_root() is not None


def _preproc() -> str:
    root = source_root()
    if root is None:
        raise FileNotFoundError("neither /root/reference nor oracle/_ref is present (run oracle/build_ref.py where the "
                                "reference tree exists)")
    return os.path.join(root, "src", "preprocessing")


def load_sampler_fns():
    """(sample_representative_frames, sample_frames_uniform): the reference's own module, executed unmodified."""
    path = os.path.join(_preproc(), "datautils", "utils.py")
    spec = importlib.util.spec_from_file_location("sasvqa_reference_datautils_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.sample_representative_frames, mod.sample_frames_uniform


def load_sample_frame_indices():
    """``sample_frame_indices`` lifted from extract_features.py without importing the module."""
    import numpy as np
    path = os.path.join(_preproc(), "extract_features.py")
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "sample_frame_indices":
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns["sample_frame_indices"]
    raise LookupError("sample_frame_indices not found in " + path)
