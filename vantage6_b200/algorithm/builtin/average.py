"""Federated column average -- vantage6's canonical example algorithm:
``master`` creates an ``average_partial`` sub-task for every organization, waits, and combines
``sum_i / count_i`` into the global mean."""

from ._common import collect


def master(client, data, column_name: str, organization_ids=None):
    orgs = client.get_organizations_in_my_collaboration()
    ids = organization_ids or [o.get("id") for o in orgs]
    task = client.create_new_task(
        input_={"method": "average_partial", "kwargs": {"column_name": column_name}}, organization_ids=ids)
    results = collect(client, task, ids)          # every node, or an error: never a mean over fewer nodes than asked for
    global_sum = sum(float(r["sum"]) for r in results)
    global_count = sum(int(r["count"]) for r in results)
    return {"average": global_sum / global_count, "count": global_count}


def RPC_average_partial(data, column_name: str):
    col = data[column_name]
    return {"sum": float(col.sum()), "count": int(len(col))}
