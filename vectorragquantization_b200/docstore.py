"""Stand-in for ``rocksdict.Rdict`` (not installable here; storage engine is out of scope - DESIGN.md section 8).

The reference uses exactly this slice of Rdict (CohereEnhancedVectorDB.py:88,191,221,303,335; VectorDBInt8.py:37,
158,179,228,251): ``Rdict(path, options)``, ``key in db``, ``db[key] = value``, ``db.get(key[, default])``,
``del db[key]``.  Values are small dicts holding the document text; the numeric payloads the reference keeps in the
same pickles live in the device-resident index instead (binary_index.py), because that is what the GPU rescoring
kernels read.  Persistence: an append-only pickle log under ``<path>/docs.log`` replayed on open.
"""
from __future__ import annotations

import os
import pickle
from typing import Any, Dict, Optional


class DocStore:
    def __init__(self, path: Optional[str] = None, options: Any = None):
        self.path = path
        self._d: Dict[str, Any] = {}
        self._fh = None
        if path is not None:
            os.makedirs(path, exist_ok=True)
            log = os.path.join(path, "docs.log")
            if os.path.exists(log):
                with open(log, "rb") as f:
                    while True:
                        try:
                            op, k, v = pickle.load(f)
                        except EOFError:
                            break
                        if op == "set":
                            self._d[k] = v
                        else:
                            self._d.pop(k, None)
            self._fh = open(log, "ab")

    def _log(self, rec):
        if self._fh is not None:
            pickle.dump(rec, self._fh, protocol=4)
            self._fh.flush()

    def __contains__(self, key: str) -> bool:
        return key in self._d

    def __setitem__(self, key: str, value: Any) -> None:
        self._d[key] = value
        self._log(("set", key, value))

    def __getitem__(self, key: str) -> Any:
        return self._d[key]

    def __delitem__(self, key: str) -> None:
        del self._d[key]
        self._log(("del", key, None))

    def get(self, key: str, default: Any = None) -> Any:
        return self._d.get(key, default)

    def __len__(self) -> int:
        return len(self._d)

    def keys(self):
        return self._d.keys()

    def close(self) -> None:
        if self._fh is not None:
            self._fh.close()
            self._fh = None


Rdict = DocStore  # the name the reference imports (``from rocksdict import Rdict``)
