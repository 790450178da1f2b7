"""Federated contingency table (the ``v6-crosstab-py`` class of algorithm): counts of ``row`` x ``column`` levels over
every node's data, with the chi-square statistic of independence computed by the master from the pooled table.

One round; a node reports its local table with cells below ``min_count`` (default 5) zeroed and flagged, and nothing at
all when it holds fewer than ``min_rows`` rows.
"""
from ._common import collect, effective_min_count, guard_rows

MIN_ROWS, MIN_COUNT = 10, 5


def master(client, data, row: str, column: str, organization_ids=None, min_rows: int = MIN_ROWS, min_count: int = MIN_COUNT):
    ids = organization_ids or [o.get("id") for o in client.get_organizations_in_my_collaboration()]
    kw = {"row": row, "column": column, "min_rows": min_rows, "min_count": min_count}
    t = client.create_new_task(input_={"method": "crosstab_partial", "kwargs": kw}, organization_ids=ids)
    parts = collect(client, t, ids)
    table: dict = {}
    for p in parts:
        for r, cells in p["table"].items():
            for c, n in cells.items():
                table.setdefault(r, {})[c] = table.get(r, {}).get(c, 0) + int(n)
    rows = sorted(table)
    cols = sorted({c for cells in table.values() for c in cells})
    dense = [[table[r].get(c, 0) for c in cols] for r in rows]
    total = sum(map(sum, dense))
    chi2, dof = None, (len(rows) - 1) * (len(cols) - 1)
    if total and dof > 0:
        rs, cs = [sum(x) for x in dense], [sum(x[j] for x in dense) for j in range(len(cols))]
        chi2 = sum((dense[i][j] - rs[i] * cs[j] / total) ** 2 / (rs[i] * cs[j] / total)
                   for i in range(len(rows)) for j in range(len(cols)) if rs[i] and cs[j])
    return {"rows": rows, "columns": cols, "table": dense, "n": total, "chi2": chi2, "dof": dof,
            "suppressed": any(p["suppressed"] for p in parts), "n_nodes": len(parts)}


def RPC_crosstab_partial(data, row: str, column: str, min_rows: int = MIN_ROWS, min_count: int = MIN_COUNT):
    guard_rows(len(data), min_rows, "report counts")
    min_count = effective_min_count(min_count)
    sub = data[[row, column]].dropna().astype(str)
    counts = sub.groupby([row, column]).size()
    table, suppressed = {}, False
    for (r, c), n in counts.items():
        if n < min_count:
            n, suppressed = 0, True
        table.setdefault(r, {})[c] = int(n)
    return {"table": table, "suppressed": suppressed}
