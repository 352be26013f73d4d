"""Importable alias of the package directory ``aes-gcm-128-192-256-bits_b200/``
(whose name, fixed by the project layout, is not a Python identifier): this
module's search path is that directory, so ``aesgcm_b200.engine`` etc. resolve
to the files there."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                          "aes-gcm-128-192-256-bits_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f
