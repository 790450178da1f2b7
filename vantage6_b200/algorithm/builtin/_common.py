"""Shared by the tabular algorithms."""


def collect(client, task, ids):
    """Wait for a sub-task and return its partial results -- all of them: a node that refused (privacy guard) or failed
    must stop the analysis, not silently shrink it."""
    parts = client.wait_for_results(task.get("id"))
    if len(parts) != len(ids):
        raise RuntimeError(f"{len(ids) - len(parts)} of {len(ids)} nodes returned no result (refused or failed: see the logs of their results)")
    return parts


# ---------------------------------------------------------------------------------------------------------------------
# The data station's floors.  ``min_rows`` / ``min_count`` arrive in the task input, i.e. from the researcher: they may ask
# for MORE protection than the node's floor, never for less.  The floor is the node operator's (`privacy: {min_rows,
# min_count}` in the node configuration, exported as V6B200_MIN_ROWS / V6B200_MIN_COUNT), 10 rows / 5 per cell by default.
DEFAULT_MIN_ROWS, DEFAULT_MIN_COUNT = 10, 5


def _floor(env: str, default: int) -> int:
    import os

    try:
        return max(0, int(os.environ.get(env, default)))
    except ValueError:
        return default


def effective_min_rows(requested=None) -> int:
    floor = _floor("V6B200_MIN_ROWS", DEFAULT_MIN_ROWS)
    return floor if requested is None else max(floor, int(requested))


def effective_min_count(requested=None) -> int:
    floor = _floor("V6B200_MIN_COUNT", DEFAULT_MIN_COUNT)
    return floor if requested is None else max(floor, int(requested))


def guard_rows(n_rows: int, requested, what: str) -> None:
    """Refuse to answer from fewer rows than the node's floor (or the stricter number the task asked for)."""
    need = effective_min_rows(requested)
    if n_rows < need:
        raise PermissionError(f"this node holds fewer than {need} rows: refusing to {what}")

