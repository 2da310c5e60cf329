"""Import the reference (p768lwy3/torecsys) hot-path sub-packages in the BUILD CONTAINER only.

TEST INFRASTRUCTURE.  `/root/reference` does not exist on the GPU box, so nothing that runs
there (`-m gpu` tests, smoke(), bench.py) may call `load_reference()`; only
`oracle/make_golden.py` and the container-only conformance tests do (they skip when absent).

Why a shim: `torecsys/__init__.py:7-15` eagerly imports cli/data/trainer, which need
pytorch_lightning, texttable, torchmetrics, ... (absent, no network), and
`torecsys/utils/operations.py:9-10` imports matplotlib at module top.  We register an empty
package object whose __path__ points at the reference tree, stub matplotlib, and import only
`torecsys.inputs`, `torecsys.layers`, `torecsys.models`.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pick_root() -> str:
    """TORECSYS_REFERENCE, else the source tree of the build container, else the UNMODIFIED reference installed by
    `pip install --no-deps --target baseline/_ref` (git-ignored; it travels to the GPU box, where bench.py's
    reference arm and cpu_baseline use it -- the tests never do)."""
    cands = [os.environ.get('TORECSYS_REFERENCE'), '/root/reference', os.path.join(_REPO, 'baseline', '_ref')]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, 'torecsys', 'layers')):
            return c
    return cands[1]


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'torecsys', 'layers'))


def load_reference():
    """Returns the `torecsys` package object (inputs, layers, models imported)."""
    if not reference_available():
        raise RuntimeError(f'reference tree not found under {REFERENCE_ROOT}')
    if 'torecsys' in sys.modules and getattr(sys.modules['torecsys'], '__b200_shim__', False):
        return sys.modules['torecsys']
    pkg = types.ModuleType('torecsys')
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, 'torecsys')]
    pkg.__b200_shim__ = True
    sys.modules['torecsys'] = pkg
    for name in ('matplotlib', 'matplotlib.pyplot', 'matplotlib.ticker'):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torecsys.inputs  # noqa: F401
    import torecsys.layers  # noqa: F401
    import torecsys.models  # noqa: F401
    from torecsys.models.sequential import Sequential
    pkg.Sequential = Sequential
    return pkg
