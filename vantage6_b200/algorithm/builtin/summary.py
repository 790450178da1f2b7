"""Federated summary statistics (the ``v6-summary-py`` class of algorithm): per numeric column the global count, mean,
standard deviation, minimum, maximum and number of missing values; per categorical column the level counts.

Two rounds on the control plane, a few floats per column each way: round one returns count / sum / min / max / missing
(and the level counts), the master forms the global means and hands them back; round two returns the sums of squared
deviations from the GLOBAL mean, so the pooled variance is exact rather than an average of node variances.

Privacy guard (the node's side of the bargain): a node refuses to answer when it holds fewer than ``min_rows`` rows
(default 10), and level counts below ``min_count`` are reported as 0 with ``suppressed: true``.
"""
import math

import numpy as np

from ._common import collect, effective_min_count, guard_rows

MIN_ROWS, MIN_COUNT = 10, 5


def _columns(data, columns):
    cols = list(columns) if columns else list(data.columns)
    missing = [c for c in cols if c not in data.columns]
    if missing:
        raise KeyError(f"column(s) not in this node's data: {missing}")
    return cols


def _numeric(series) -> bool:
    return series.dtype.kind in "biuf"


def master(client, data, columns=None, organization_ids=None, min_rows: int = MIN_ROWS, min_count: int = MIN_COUNT,
           quantiles=None, bins: int = 512):
    """``quantiles`` (e.g. ``[0.25, 0.5, 0.75]``) adds a third round: every node histograms its numeric columns on a common
    grid of ``bins`` cells between the pooled minimum and maximum, the master reads the quantiles off the pooled histogram
    (linear within a cell: the answer is within one cell width, (max - min) / bins, of a value whose pooled rank is q)."""
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    kw = {"columns": columns, "min_rows": min_rows, "min_count": min_count}
    t = client.create_new_task(input_={"method": "summary_partial", "kwargs": kw}, organization_ids=ids)
    parts = collect(client, t, ids)
    numeric = sorted(set().union(*[set(p["numeric"]) for p in parts]))
    out, means = {}, {}
    for c in numeric:
        ps = [p["numeric"][c] for p in parts if c in p["numeric"]]
        n = sum(p["count"] for p in ps)
        means[c] = sum(p["sum"] for p in ps) / n if n else float("nan")
        out[c] = {"count": n, "missing": sum(p["missing"] for p in ps), "mean": means[c],
                  "min": min(p["min"] for p in ps if p["count"]) if n else None,
                  "max": max(p["max"] for p in ps if p["count"]) if n else None}
    t2 = client.create_new_task(input_={"method": "deviation_partial", "kwargs": {"means": means, "min_rows": min_rows}},
                                organization_ids=ids)
    devs = collect(client, t2, ids)
    for c in numeric:
        ss = sum(p.get(c, 0.0) for p in devs)
        n = out[c]["count"]
        out[c]["std"] = math.sqrt(ss / (n - 1)) if n > 1 else float("nan")
    if quantiles:
        qs = [float(q) for q in quantiles]
        if any(not 0.0 <= q <= 1.0 for q in qs):
            raise ValueError("quantiles must lie in [0, 1]")
        ranges = {c: (out[c]["min"], out[c]["max"]) for c in numeric if out[c]["count"]}
        t3 = client.create_new_task(input_={"method": "histogram_partial", "kwargs": {"ranges": ranges, "bins": bins, "min_rows": min_rows}},
                                    organization_ids=ids)
        hists = collect(client, t3, ids)
        for c, (lo, hi) in ranges.items():
            counts = [sum(h[c][i] for h in hists) for i in range(bins)]
            out[c]["quantiles"] = {str(q): _quantile(counts, lo, hi, q) for q in qs}
    for c in sorted(set().union(*[set(p["categorical"]) for p in parts])):
        levels, suppressed = {}, False
        for p in parts:
            pc = p["categorical"].get(c)
            if pc is None:
                continue
            suppressed |= bool(pc["suppressed"])
            for k, v in pc["counts"].items():
                levels[k] = levels.get(k, 0) + int(v)
        out[c] = {"counts": levels, "suppressed": suppressed}
    return {"n_rows": sum(p["n_rows"] for p in parts), "n_nodes": len(parts), "columns": out}


def _quantile(counts, lo: float, hi: float, q: float) -> float:
    total = sum(counts)
    if total == 0 or hi <= lo:
        return lo
    target, width, acc = q * total, (hi - lo) / len(counts), 0.0
    for i, n in enumerate(counts):
        if n and acc + n >= target:
            return lo + (i + (target - acc) / n) * width
        acc += n
    return hi


def RPC_histogram_partial(data, ranges: dict, bins: int = 512, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "report statistics")
    out = {}
    for c, (lo, hi) in ranges.items():
        v = data[c].dropna().to_numpy(dtype=float)
        if hi <= lo:
            out[c] = [int(len(v))] + [0] * (bins - 1)
            continue
        idx = ((v - lo) / (hi - lo) * bins).astype(int).clip(0, bins - 1)
        out[c] = [int(n) for n in np.bincount(idx, minlength=bins)]
    return out


def RPC_summary_partial(data, columns=None, min_rows: int = MIN_ROWS, min_count: int = MIN_COUNT):
    guard_rows(len(data), min_rows, "report statistics")
    min_count = effective_min_count(min_count)
    numeric, categorical = {}, {}
    for c in _columns(data, columns):
        s = data[c]
        if _numeric(s):
            v = s.dropna()
            numeric[c] = {"count": int(len(v)), "missing": int(s.isna().sum()), "sum": float(v.sum()),
                          "min": float(v.min()) if len(v) else None, "max": float(v.max()) if len(v) else None}
        else:
            counts = s.dropna().astype(str).value_counts()
            small = counts < min_count
            categorical[c] = {"counts": {k: (0 if small[k] else int(n)) for k, n in counts.items()}, "suppressed": bool(small.any())}
    return {"n_rows": int(len(data)), "numeric": numeric, "categorical": categorical}


def RPC_deviation_partial(data, means: dict, min_rows: int = MIN_ROWS):
    guard_rows(len(data), min_rows, "report statistics")
    return {c: float(((data[c].dropna() - m) ** 2).sum()) for c, m in means.items() if c in data.columns}
