"""Shared by the tabular algorithms."""


def collect(client, task, ids):
    """Wait for a sub-task and return its partial results -- all of them: a node that refused (privacy guard) or failed
    must stop the analysis, not silently shrink it."""
    parts = client.wait_for_results(task.get("id"))
    if len(parts) != len(ids):
        raise RuntimeError(f"{len(ids) - len(parts)} of {len(ids)} nodes returned no result (refused or failed: see the logs of their results)")
    return parts
