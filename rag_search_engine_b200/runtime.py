"""One librse handle per (database file, device), shared by the mirror classes that sit on it.

The reference opens two SQLite connections per HybridSearch (hybrid_search.py:41-54); here the
keyword and semantic halves share ONE device handle so the fused on-device hybrid path
(``rse_hybrid``) can see both indexes.  Reference counted; the last ``close()`` frees HBM.
"""
from __future__ import annotations

from pathlib import Path
from typing import Dict, Tuple

from . import _lib

_handles: Dict[Tuple[str, int], list] = {}


def acquire(db_path, device: int = 0) -> "_lib.Index":
    key = (str(Path(db_path).resolve()), int(device))
    ent = _handles.get(key)
    if ent is None:
        ent = [_lib.Index(device), 0, {}]          # handle, refcount, loaded-parts registry
        _handles[key] = ent
    ent[1] += 1
    return ent[0]


def adopt(key_path, device: int, index: "_lib.Index") -> None:
    """Register a handle the caller created and loaded itself (HybridSearch.from_loaded); never closed by release()."""
    _handles.setdefault((str(Path(key_path).resolve()), int(device)), [index, 1 << 30, {}])


def parts(db_path, device: int = 0) -> dict:
    return _handles[(str(Path(db_path).resolve()), int(device))][2]


def release(db_path, device: int = 0) -> None:
    key = (str(Path(db_path).resolve()), int(device))
    ent = _handles.get(key)
    if ent is None:
        return
    ent[1] -= 1
    if ent[1] <= 0:
        ent[0].close()
        del _handles[key]
