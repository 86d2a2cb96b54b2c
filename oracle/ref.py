"""oracle/ref.py -- TEST INFRASTRUCTURE.  Loader for the reference's own compiled Cython modules
(``oracle/_ref/*.so`` built by ``oracle/build_ref.py``).  Exposes the reference functions unmodified:
``process_block_nonzero``, ``kernel``, ``find_object_properties``, ``map_subcell_extract_props``,
``map_subcell_C``, ``extract_cs_syntype``.  ``available()`` is False when the modules were not built."""
import glob
import importlib.util
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")
_mods = {}


def _load(name):
    if name not in _mods:
        cands = glob.glob(os.path.join(_REF, name + "*.so"))
        if not cands:
            raise ImportError(f"oracle/_ref/{name}*.so missing: run python oracle/build_ref.py")
        spec = importlib.util.spec_from_file_location(name, cands[0])
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        _mods[name] = m
    return _mods[name]


def available():
    try:
        _load("block_processing_C")
        _load("find_object_properties_C")
        return True
    except Exception:
        return False


def __getattr__(name):
    if name in ("process_block_nonzero", "kernel", "extract_cs_syntype"):
        return getattr(_load("block_processing_C"), name)
    if name in ("find_object_properties", "map_subcell_extract_props", "map_subcell_C"):
        return getattr(_load("find_object_properties_C"), name)
    raise AttributeError(name)
