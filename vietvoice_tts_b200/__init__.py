"""Importable alias for the `vietvoice-tts_b200/` package directory.

The package directory carries the reference's hyphenated name (not a legal Python identifier);
this stub makes `import vietvoice_tts_b200` resolve into it.
"""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "vietvoice-tts_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py"), "r", encoding="utf-8") as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
