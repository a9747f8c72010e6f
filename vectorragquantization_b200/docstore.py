"""Stand-in for ``rocksdict.Rdict`` (not installable here; storage engine is out of scope - DESIGN.md section 8).

The reference uses exactly this slice of Rdict (CohereEnhancedVectorDB.py:88,191,221,303,335; VectorDBInt8.py:37,
158,179,228,251): ``Rdict(path, options)``, ``key in db``, ``db[key] = value``, ``db.get(key[, default])``,
``del db[key]``.  Values are small dicts holding the document text; the numeric payloads the reference keeps in the
same pickles live in the device-resident index instead (binary_index.py), because that is what the GPU rescoring
kernels read.

Persistence: an append-only pickle log ``<path>/docs.log`` replayed on open (a torn last record - a crash in the middle
of a write - is cut off, not fatal; a log that is mostly dead records is compacted on open).  When ``<path>`` holds a
store WRITTEN BY THE REFERENCE (RocksDB table files), its documents are imported read-only as the base the log is
replayed on top of (rocks_import.py); the imported raw values stay available in ``imported_raw`` until the owning class
has moved their vectors into the index.
"""
from __future__ import annotations

import logging
import os
import pickle
from typing import Any, Dict, Iterable, Optional, Tuple

logger = logging.getLogger(__name__)


class DocStore:
    def __init__(self, path: Optional[str] = None, options: Any = None):
        self.path = path
        self._d: Dict[str, Any] = {}
        self._fh = None
        self.imported_raw: Optional[Dict[str, Any]] = None
        if path is None:
            return
        os.makedirs(path, exist_ok=True)
        from . import rocks_import
        if rocks_import.is_rocksdict_folder(path):
            raw = rocks_import.read_rocksdict_folder(path)
            self.imported_raw = {str(k): v for k, v in raw.items()}
            for k, v in self.imported_raw.items():
                self._d[k] = {"doc": v.get("doc", "N/A")} if isinstance(v, dict) else v
            logger.info("Imported %d documents from the RocksDB store in %s (read-only).", len(self._d), path)
        self._replay()

    # ---- log ------------------------------------------------------------------------------------------------
    def _log_path(self) -> str:
        return os.path.join(self.path, "docs.log")

    def _replay(self) -> None:
        log = self._log_path()
        if not os.path.exists(log):
            return
        records, good_end = 0, 0
        with open(log, "rb") as f:
            while True:
                try:
                    rec = pickle.load(f)
                    if rec[0] == "many":
                        for k, v in rec[1]:
                            self._d[k] = v
                    elif rec[0] == "set":
                        self._d[rec[1]] = rec[2]
                    else:
                        self._d.pop(rec[1], None)
                    records += len(rec[1]) if rec[0] == "many" else 1
                    good_end = f.tell()
                except EOFError:
                    break
                except Exception as e:  # torn / corrupt tail: keep what was read, drop the rest
                    logger.warning("docs.log: dropping a damaged tail after %d records (%s)", records, e)
                    break
        size = os.path.getsize(log)
        try:
            if good_end < size:
                with open(log, "r+b") as f:
                    f.truncate(good_end)
            if records > 2 * len(self._d) + 1024:
                self._compact()
        except OSError:  # read-only folder: the in-memory state is what matters
            pass

    def _compact(self) -> None:
        tmp = self._log_path() + ".tmp"
        base = set(self.imported_raw or ())
        with open(tmp, "wb") as f:
            pickle.dump(("many", list(self._d.items())), f, protocol=4)
            for k in base - set(self._d):
                pickle.dump(("del", k, None), f, protocol=4)
            f.flush()
            os.fsync(f.fileno())
        os.replace(tmp, self._log_path())

    def _log(self, rec) -> None:
        if self.path is None:
            return
        if self._fh is None:
            self._fh = open(self._log_path(), "ab")
        pickle.dump(rec, self._fh, protocol=4)
        self._fh.flush()

    # ---- the Rdict slice the reference uses ---------------------------------------------------------------------
    def __contains__(self, key: str) -> bool:
        return key in self._d

    def __setitem__(self, key: str, value: Any) -> None:
        self._d[key] = value
        self._log(("set", key, value))

    def set_many(self, items: Iterable[Tuple[str, Any]]) -> None:
        """Bulk path: one log record and one flush for a whole batch (the reference writes one pickle per document)."""
        items = list(items)
        if not items:
            return
        self._d.update(items)
        self._log(("many", items))

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
