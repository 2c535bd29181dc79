"""Loads the reference's own functions when /root/reference is present (build container
only -- it does not exist on the GPU box).  TEST INFRASTRUCTURE ONLY.

``datautils/utils.py`` imports cleanly (torch + heapq only).  ``extract_features.py`` does
not (h5py/cv2 missing), so ``sample_frame_indices`` (``extract_features.py:32-39``) is
compiled at run time from that file's own AST node -- nothing is copied into this repo.
"""
from __future__ import annotations

import ast
import os
import sys

REFERENCE_ROOT = os.environ.get("SASVQA_REFERENCE_ROOT", "/root/reference")
_PREPROC = os.path.join(REFERENCE_ROOT, "src", "preprocessing")


def available() -> bool:
    return os.path.isfile(os.path.join(_PREPROC, "datautils", "utils.py"))


def load_sampler_fns():
    """(sample_representative_frames, sample_frames_uniform) from the reference tree."""
    if not available():
        raise FileNotFoundError(_PREPROC)
    if _PREPROC not in sys.path:
        sys.path.insert(0, _PREPROC)
    from datautils.utils import sample_representative_frames, sample_frames_uniform  # type: ignore
    return sample_representative_frames, sample_frames_uniform


def load_sample_frame_indices():
    """``sample_frame_indices`` lifted from extract_features.py without importing the module."""
    import numpy as np
    path = os.path.join(_PREPROC, "extract_features.py")
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == "sample_frame_indices":
            ns = {"np": np}
            exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
            return ns["sample_frame_indices"]
    raise LookupError("sample_frame_indices not found in " + path)
