"""Payload (de)serialisation for task inputs and results on the control plane.

vantage6 3.x ships inputs/results as JSON (default) or pickle (legacy) blobs; algorithm outputs
here are frequently numeric arrays, so the JSON flavour understands numpy arrays and torch
tensors (encoded as ``{"__ndarray__": b64, "dtype", "shape"}``).  Large tensors should not use
this path at all -- they stay in NVLink symmetric memory (parallel/symm.py).
"""
from __future__ import annotations

import base64
import json
import pickle
import sys
from typing import Any

import numpy as np


def _default(o: Any):
    torch = sys.modules.get("torch")        # never import torch here: it costs seconds per algorithm run
    if torch is not None and isinstance(o, torch.Tensor):
        t = o.detach().cpu()
        if t.dtype == torch.bfloat16:
            t = t.float()
        o = t.numpy()
    if isinstance(o, np.ndarray):
        a = np.asarray(o, order="C")               # (ascontiguousarray would turn a 0-d array into shape (1,))
        return {"__ndarray__": base64.b64encode(a.tobytes()).decode("ascii"), "dtype": str(a.dtype), "shape": list(a.shape)}
    if isinstance(o, (np.integer,)):
        return int(o)
    if isinstance(o, (np.floating,)):
        return float(o)
    if isinstance(o, (np.bool_,)):
        return bool(o)
    if isinstance(o, bytes):
        return {"__bytes__": base64.b64encode(o).decode("ascii")}
    if isinstance(o, (set, tuple)):
        return list(o)
    raise TypeError(f"Object of type {type(o).__name__} is not JSON serializable")


def _hook(d: dict):
    if "__ndarray__" in d:
        a = np.frombuffer(base64.b64decode(d["__ndarray__"]), dtype=np.dtype(d["dtype"])).reshape(d["shape"])
        return a.copy()
    if "__bytes__" in d:
        return base64.b64decode(d["__bytes__"])
    return d


def serialize(obj: Any, data_format: str = "json") -> bytes:
    if data_format == "json":
        return json.dumps(obj, default=_default).encode("utf-8")
    if data_format == "pickle":
        return pickle.dumps(obj)
    raise ValueError(f"unknown data format {data_format!r}")


class UnsafePayload(ValueError):
    """A pickle payload arrived where pickle was not explicitly enabled."""


def pickle_allowed() -> bool:
    """Pickle executes code on load.  Task inputs come from researchers and results from other organizations' nodes, so
    it is OFF unless the operator of this process opted in (``V6B200_ALLOW_PICKLE=1``, set from the ``allow_pickle``
    key of the node configuration / by the client's ``allow_pickle=True``)."""
    import os

    return os.environ.get("V6B200_ALLOW_PICKLE") == "1"


def deserialize(blob: bytes | str, data_format: str | None = None, allow_pickle: bool | None = None) -> Any:
    """JSON by default (the only format that is safe on untrusted bytes).  ``data_format="pickle"`` -- or an untagged blob
    that is not JSON -- is only honoured when pickle was explicitly allowed; otherwise :class:`UnsafePayload`."""
    if isinstance(blob, str):
        blob = blob.encode("utf-8")
    head = blob.lstrip()[:5]
    looks_json = (head[:1] in (b"{", b"[", b'"', b"-") or head[:1].isdigit() or head[:4] in (b"null", b"true") or head == b"false"
                  or head[:3] == b"NaN" or head == b"Infin")            # python's json writes non-finite floats as bare tokens
    if data_format is None:
        data_format = "json" if looks_json else "pickle"
    if data_format == "json":
        return json.loads(blob.decode("utf-8"), object_hook=_hook)
    if data_format != "pickle":
        raise ValueError(f"unknown data format {data_format!r}")
    if not (pickle_allowed() if allow_pickle is None else allow_pickle):
        raise UnsafePayload("refusing to unpickle a payload: pickle executes code on load and is disabled "
                            "(set allow_pickle in the node configuration / V6B200_ALLOW_PICKLE=1 to opt in)")
    return pickle.loads(blob)
