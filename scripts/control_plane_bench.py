"""Task dispatch / result collection latency through the full control plane on CPU (BASELINE config 1: weighted mean
of a 1k-parameter vector, 2 CPU nodes + 1 server): `vserver` + 2 x `vnode` processes, tasks created by a researcher
client, algorithms run by the nodes.  Prints one JSON line.

    python scripts/control_plane_bench.py                  # warm start (zygote) on
    V6B200_ZYGOTE=0 python scripts/control_plane_bench.py  # one interpreter per task
"""
import json
import os
import statistics
import tempfile
import time

import numpy as np

from vantage6_b200.dev import DemoNetwork


def main():
    home = tempfile.mkdtemp(prefix="v6cp")
    rng = np.random.default_rng(0)
    dbs = []
    for i, n in enumerate((30, 70)):
        path = f"{home}/vec{i}.npy"
        np.save(path, rng.normal(size=(n, 1000)))
        dbs.append(path)
    net = DemoNetwork(2, home=home, databases=dbs)
    net.start()
    try:
        client = net.client()
        time.sleep(3.0)                      # let the nodes' warm-start helpers come up

        def timed(orgs, input_, n):
            out = []
            for _ in range(n):
                t0 = time.perf_counter()
                task = client.task.create(collaboration=net.collaboration_id, organizations=orgs, name="t",
                                          image="v6b200/weighted-mean", input=input_)
                res = client.wait_for_results(task["id"], timeout=120)
                out.append(time.perf_counter() - t0)
                assert all(r["result"] is not None for r in res)
            return out

        master = timed([net.org_ids[0]], {"method": "master", "master": True}, 8)
        partial = timed(net.org_ids, {"method": "partial_sum"}, 8)
        print(json.dumps({"config": "weighted mean of a 1k vector, 2 CPU nodes + 1 server (BASELINE config 1)",
                          "zygote": os.environ.get("V6B200_ZYGOTE", "1") != "0",
                          "master_task_s_median": round(statistics.median(master), 4),
                          "master_task_s": [round(t, 4) for t in master],
                          "partial_task_2_nodes_s_median": round(statistics.median(partial), 4),
                          "partial_task_2_nodes_s": [round(t, 4) for t in partial]}))
    finally:
        net.stop()


if __name__ == "__main__":
    main()
