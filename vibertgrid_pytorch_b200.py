"""Import shim: makes the directory ``vibertgrid-pytorch_b200/`` (not a legal
Python identifier) importable as the package ``vibertgrid_pytorch_b200``.

A module that defines ``__path__`` is a package, so submodules resolve inside
the hyphenated directory.  The package body itself lives in that directory's
``__init__.py`` and is executed here.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vibertgrid-pytorch_b200")]
__file__ = _os.path.join(__path__[0], "__init__.py")
with open(__file__, "r") as _f:
    exec(compile(_f.read(), __file__, "exec"))
