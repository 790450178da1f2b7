"""Federated column average -- vantage6's canonical example algorithm:
``master`` creates an ``average_partial`` sub-task for every organization, waits, and combines
``sum_i / count_i`` into the global mean."""


def master(client, data, column_name: str, organization_ids=None):
    orgs = client.get_organizations_in_my_collaboration()
    ids = organization_ids or [o.get("id") for o in orgs]
    task = client.create_new_task(
        input_={"method": "average_partial", "kwargs": {"column_name": column_name}}, organization_ids=ids)
    task_id = task.get("id")
    client.wait_for_task(task_id)
    results = client.get_results(task_id=task_id)
    global_sum = sum(float(r["sum"]) for r in results)
    global_count = sum(int(r["count"]) for r in results)
    return {"average": global_sum / global_count, "count": global_count}


def RPC_average_partial(data, column_name: str):
    col = data[column_name]
    return {"sum": float(col.sum()), "count": int(len(col))}
