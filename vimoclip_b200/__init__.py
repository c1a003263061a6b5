"""Import alias: ``import vimoclip_b200`` resolves to the package directory ``vimo-clip_b200/``.

The repository layout names the package ``vimo-clip_b200/`` (after the reference repository), which
is not a valid Python identifier; this shim executes that directory's ``__init__`` under the
importable name so ``vimoclip_b200.student`` etc. resolve to files in ``vimo-clip_b200/``.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vimo-clip_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
del _os, _f
