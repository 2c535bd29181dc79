"""Import shims that let the reference's consumer modules (oracle/_ref/src/datasets/*) import in this image, which lacks
h5py, easydict and tensorboardX.  TEST INFRASTRUCTURE ONLY -- installed into ``sys.modules`` by tests/test_consumer_ref.py (pytest's rootdir conftest puts tests/ on sys.path).

* ``h5py``: ``File(path, 'r')[name]`` backed by the repo's minimal HDF5 reader (``sasvqa_b200.hdf5_min``), the two calls
  the reference makes (src/datasets/dataset_base.py:104).  When the real h5py is importable it is used instead.
* ``easydict.EasyDict`` / ``tensorboardX.SummaryWriter``: imported at module level by src/utils/{load_save,logger}.py,
  never reached by the dataset / collator code under test.
"""
from __future__ import annotations

import sys
import types


def _h5py_shim():
    from sasvqa_b200 import hdf5_min

    class File:
        def __init__(self, path, mode="r"):
            if mode not in ("r", "r+"):
                raise NotImplementedError("the h5py shim only reads (the writer under test is sasvqa_b200.writer)")
            self._sets = hdf5_min.open_datasets(path, mode)

        def __getitem__(self, name):
            return self._sets[name]

        def keys(self):
            return self._sets.keys()

        def close(self):
            self._sets = {}

        def __enter__(self):
            return self

        def __exit__(self, *exc):
            self.close()

    mod = types.ModuleType("h5py")
    mod.File = File
    mod.__shim__ = "sasvqa_b200.hdf5_min"
    return mod


def install() -> dict:
    """Installs whichever shims are needed; returns {module name: 'real' | 'shim'}."""
    used = {}
    try:
        import h5py  # noqa: F401
        used["h5py"] = "real"
    except Exception:  # noqa: BLE001
        sys.modules["h5py"] = _h5py_shim()
        used["h5py"] = "shim"
    try:
        import easydict  # noqa: F401
        used["easydict"] = "real"
    except Exception:  # noqa: BLE001
        mod = types.ModuleType("easydict")

        class EasyDict(dict):
            __getattr__ = dict.get
            __setattr__ = dict.__setitem__

        mod.EasyDict = EasyDict
        sys.modules["easydict"] = mod
        used["easydict"] = "shim"
    try:
        import tensorboardX  # noqa: F401
        used["tensorboardX"] = "real"
    except Exception:  # noqa: BLE001
        mod = types.ModuleType("tensorboardX")

        class SummaryWriter:
            def __init__(self, *a, **k):
                pass

        mod.SummaryWriter = SummaryWriter
        sys.modules["tensorboardX"] = mod
        used["tensorboardX"] = "shim"
    return used
