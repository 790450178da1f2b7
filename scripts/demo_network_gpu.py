"""Full-stack run on the GPU box: ``vserver import/start`` + N x ``vnode start --gpu k`` (process runtime), a
researcher client, and the FedAvg algorithm dispatched through the control plane; the train partials rendezvous
and aggregate over the NVLink symmetric heap (data plane "native").  Prints one JSON line per model.

    python scripts/demo_network_gpu.py --nodes 8 --model resnet50 --rounds 4
"""
import argparse
import json
import tempfile
import time

from vantage6_b200.dev import DemoNetwork


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=2)
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--rounds", type=int, default=4)
    ap.add_argument("--glm", action="store_true", help="also run the fused GLM (K8 + K3) task")
    ap.add_argument("--repeat", type=int, default=2, help="FedAvg tasks on the same network: the 2nd.. reuse the resident GPU workers")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    home = tempfile.mkdtemp(prefix="v6gpu")
    net = DemoNetwork(args.nodes, home=home, gpus=list(range(args.nodes)))
    lines = []
    try:
        net.start(timeout=180)
        c = net.client()
        jobs = [("v6b200/fedavg", {"method": "master", "master": True,
                                   "kwargs": {"model": args.model, "rounds": args.rounds, "return_weights": True, "seed": 0}})
                for _ in range(max(1, args.repeat))]
        if args.glm:
            jobs.append(("v6b200/glm", {"method": "master_fused", "master": True,
                                        "kwargs": {"iterations": 200, "lr": 2.0, "rows_per_node": 125000, "features": 256,
                                                   "synthetic": True}}))
        for image, inp in jobs:
            t0 = time.time()
            task = c.task.create(collaboration=net.collaboration_id, organizations=[net.org_ids[0]], name=image, image=image,
                                 input=inp)
            try:
                res = c.wait_for_results(task["id"], timeout=240)
            except TimeoutError:
                print(net.tail_logs(60))
                raise
            out = res[0]["result"]
            if out is None:
                print(res[0]["log"])
                raise SystemExit(1)
            out = {k: (v.tolist() if hasattr(v, "tolist") else v) for k, v in out.items() if k not in ("coef", "weights", "w")}
            out = {k: v for k, v in out.items() if not (isinstance(v, list) and len(v) > 64)}
            out.update({"image": image, "nodes": args.nodes, "task_wall_s": round(time.time() - t0, 2),
                        "task_index": len(lines),
                        "trainer_reused": [n.get("trainer_reused") for n in out.get("nodes", [])] if isinstance(out.get("nodes"), list) else None,
                        "setup_s": [round(n.get("setup_s") or 0.0, 3) for n in out.get("nodes", [])] if isinstance(out.get("nodes"), list) else None})
            lines.append(out)
            print(json.dumps(out), flush=True)
            if args.out:                                    # incremental: a later task may time out
                with open(args.out, "a") as f:
                    f.write(json.dumps(out) + "\n")
    finally:
        net.stop()


if __name__ == "__main__":
    main()
