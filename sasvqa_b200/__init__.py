"""Import alias: the package lives in ``sas-vqa_b200/`` (not a valid Python identifier),
this shim makes it importable as ``sasvqa_b200`` by pointing ``__path__`` there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "sas-vqa_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
