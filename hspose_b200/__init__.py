"""Importable alias of the product package directory `hs-pose_b200/`.

The directory name mandated for the package contains a hyphen and cannot be
written in an `import` statement; this shim makes `import hspose_b200` (and
`hspose_b200.<submodule>`) resolve into that directory.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "hs-pose_b200")
__path__ = [_real]
__file__ = _os.path.join(_real, "__init__.py")
with open(__file__) as _f:
    exec(compile(_f.read(), __file__, "exec"))
