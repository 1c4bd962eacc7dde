"""Recipe for ``oracle/_ref``: the UNMODIFIED reference, byte-compiled from where it lies.

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as oracle/cough_oracle.py: nothing under
``cough_detector_b200/`` may import this).  ``/root/reference`` exists only in the authoring
container, so the reference cannot be imported on the GPU box.  What can travel is a BUILT
artefact, exactly like a compiled C reference would: ``build()`` runs CPython's own compiler
over ``/root/reference/src/preprocessing.py`` and ``src/model.py`` (the two files on the path,
SURVEY.md §8a; neither imports anything else from the reference) and writes sourceless
bytecode to ``oracle/_ref/*.bytecode`` (``.pyc`` content).  No reference source text is copied; ``oracle/_ref/`` is
git-ignored and not gpurun-ignored, so the bytecode travels with the snapshot to a box that
runs this same interpreter.

``load()`` returns the two modules (or ``None`` when ``oracle/_ref`` has not been built or was
built by another interpreter version); ``bench.py``'s CPU legs then time the reference itself
(``cpu_baseline.kind = "reference"``) and fall back to ``oracle/reference_port.py``
(``"port"``) otherwise.  ``tests/test_oracle.py`` checks the port against it bit for bit.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import py_compile
import sys
from types import ModuleType
from typing import Dict, Optional

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"
OUT = os.path.join(HERE, "_ref")
FILES = ("preprocessing", "model")          # src/preprocessing.py, src/model.py
EXT = ".bytecode"                           # CPython .pyc content; not named *.pyc so no sync tool filters it out
_cache: Optional[Dict[str, ModuleType]] = None
last_error = ""                             # why load() returned None (bench.py prints it on stderr)


def build(verbose: bool = True) -> bool:
    """Compile the reference's two files into oracle/_ref/.  Returns False (and leaves any existing output alone)
    when /root/reference is absent, as on the GPU box."""
    if not all(os.path.isfile(os.path.join(REF_SRC, f + ".py")) for f in FILES):
        return False
    os.makedirs(OUT, exist_ok=True)
    for f in FILES:
        py_compile.compile(os.path.join(REF_SRC, f + ".py"), cfile=os.path.join(OUT, f + EXT),
                           dfile=f"<reference>/src/{f}.py", doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(os.path.join(OUT, "BUILT_WITH"), "w") as fh:
        fh.write(sys.version.split()[0] + " " + importlib.util.MAGIC_NUMBER.hex() + "\n")
    if verbose:
        print("oracle/_ref: compiled", ", ".join(f"src/{f}.py" for f in FILES), "from", REF_SRC)
    return True


def _load_one(name: str) -> ModuleType:
    path = os.path.join(OUT, name + EXT)
    loader = importlib.machinery.SourcelessFileLoader("cough_reference_" + name, path)
    spec = importlib.util.spec_from_loader(loader.name, loader)
    mod = importlib.util.module_from_spec(spec)
    loader.exec_module(mod)
    return mod


def load() -> Optional[Dict[str, ModuleType]]:
    """{'preprocessing': module, 'model': module} of the unmodified reference, or None."""
    global _cache, last_error
    if _cache is not None:
        return _cache
    try:
        with open(os.path.join(OUT, "BUILT_WITH")) as fh:
            magic = fh.read().split()[1]
        if magic != importlib.util.MAGIC_NUMBER.hex():
            last_error = f"oracle/_ref was built by another interpreter (magic {magic})"
            return None
        _cache = {f: _load_one(f) for f in FILES}
    except (OSError, IndexError, ImportError) as e:
        last_error = f"{type(e).__name__}: {e}"
        return None
    return _cache


if __name__ == "__main__":
    ok = build()
    print("built" if ok else "no /root/reference here: nothing built", "->", OUT)
