"""Node-local training data for the federated-learning algorithms.

The node runtime hands an algorithm the database the task names (``DATABASE_URI`` / ``{LABEL}_DATABASE_URI``: the
contract of reference vantage6/cli/node.py:360-378; loaded by algorithm/wrapper.py::load_data).  This module turns that
object into the list of pinned-host ``(x, y)`` batches ``FederatedTrainer.run_round`` consumes:

* ``.npz`` / ``.pt`` shards: a mapping with the inputs under ``x`` | ``images`` | ``input_ids`` | ``features`` and the
  targets under ``y`` | ``labels`` | ``targets`` (language-model shards may omit the targets: next-token labels are the
  inputs themselves),
* CSV / parquet frames: every numeric column except ``label`` / ``y`` / ``target`` is a feature,
* ``synthetic://...`` URIs or no database at all: the model's synthetic generator, seeded per organization (non-IID),
  which is what the benchmarks use (there is no network for real datasets).

Only the node's own process ever touches these tensors; what crosses NVLink is model parameters / deltas.
"""
from __future__ import annotations

from typing import Any, List, Optional, Tuple

X_KEYS = ("x", "images", "input_ids", "features", "X")
Y_KEYS = ("y", "labels", "targets", "label", "target")


def _pick(d: dict, keys) -> Optional[Any]:
    for k in keys:
        if k in d:
            return d[k]
    return None


def is_synthetic(data: Any) -> bool:
    return data is None or (isinstance(data, str) and (data.startswith("synthetic://") or data == ""))


def tensors_from_database(data: Any):
    """-> (x, y) torch tensors on the host, or None when the database is synthetic / absent."""
    import numpy as np
    import torch

    if is_synthetic(data):
        return None
    if isinstance(data, str):
        raise ValueError(f"unsupported database URI {data!r}")
    if hasattr(data, "columns") and hasattr(data, "to_numpy"):            # pandas frame
        ycol = next((c for c in data.columns if str(c).lower() in Y_KEYS), None)
        feats = data.drop(columns=[ycol]) if ycol is not None else data
        feats = feats.select_dtypes(include="number")
        x = torch.from_numpy(np.ascontiguousarray(feats.to_numpy(dtype=np.float32)))
        y = torch.from_numpy(np.ascontiguousarray(data[ycol].to_numpy())) if ycol is not None else None
        return x, y
    if isinstance(data, (tuple, list)) and len(data) == 2:
        data = {"x": data[0], "y": data[1]}
    if isinstance(data, dict):
        x, y = _pick(data, X_KEYS), _pick(data, Y_KEYS)
        if x is None:
            raise ValueError(f"database has none of the input keys {X_KEYS}: {sorted(data)[:8]}")
        to_t = lambda a: a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a))   # noqa: E731
        return to_t(x), (to_t(y) if y is not None else None)
    if torch.is_tensor(data) or isinstance(data, np.ndarray):
        t = data if torch.is_tensor(data) else torch.from_numpy(np.ascontiguousarray(data))
        return t, None
    raise ValueError(f"do not know how to train on a database of type {type(data).__name__}")


def make_local_batches(data: Any, spec, n_steps: Optional[int], batch: int, seed: int, pin: bool, source: str = "auto"
                       ) -> Tuple[List[Tuple[Any, Any]], float, str]:
    """-> (batches, n_samples, source).  ``n_steps=None`` = one local epoch over the node's data.

    ``source``: ``"database"`` trains on the node's database or fails, ``"synthetic"`` ignores it, ``"auto"`` (default)
    uses the database when it is a labelled dataset (a mapping / frame / pair with inputs AND targets, or a token
    shard) and otherwise falls back to the synthetic generator -- a node whose default database is, say, a bare
    feature matrix for another algorithm can still take part in a benchmark run."""
    import torch

    xy = None if source == "synthetic" else tensors_from_database(data)
    if xy is not None and source == "auto":
        x, y = xy
        labelled = y is not None or (not x.dtype.is_floating_point and x.dim() == 2)      # token shard: labels = inputs
        if not labelled:
            xy = None
    if xy is None:
        if source == "database":
            raise ValueError("data_source='database' but the node has no usable labelled database")
        steps = n_steps or spec.local_steps
        return spec.make_batches(steps, batch, seed, pin=pin), float(steps * batch), "synthetic"
    x, y = xy
    n = int(x.shape[0])
    if n < batch:
        raise ValueError(f"the node's database holds {n} samples, fewer than one batch of {batch}")
    steps = n // batch if n_steps is None else min(int(n_steps), n // batch)
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm(n, generator=g)[: steps * batch]
    if y is None:
        y = x                                       # language-model shard: labels are the (shifted) inputs
    out = []
    for i in range(steps):
        idx = perm[i * batch:(i + 1) * batch]
        xb, yb = x.index_select(0, idx).contiguous(), y.index_select(0, idx).contiguous()
        if yb.dtype in (torch.int32, torch.int16, torch.uint8) and xb is not yb and yb.dim() == 1:
            yb = yb.long()
        if pin and torch.cuda.is_available():
            xb, yb = xb.pin_memory(), yb.pin_memory()
        out.append((xb, yb))
    return out, float(steps * batch), "database"
