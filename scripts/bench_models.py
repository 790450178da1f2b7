#!/usr/bin/env python
"""Secondary BASELINE.json configs (3: BERT-base bf16 4 local steps/round, 4: Llama-3 8B LoRA,
5: logistic GLM 1M x 256) for both arms, any number of GPUs (torchrun) -- same timing rules as
bench.py (CUDA events, barrier+sync on both sides, max over ranks, >=3 warm-up rounds).

    python scripts/bench_models.py --model bert_base --impl b200 --rounds 6
    torchrun --nproc-per-node 8 ... scripts/bench_models.py --model glm --impl nccl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="bert_base", choices=["bert_base", "bert_tiny", "llama3_8b_lora", "llama_tiny_lora", "glm", "resnet50"])
    ap.add_argument("--impl", default="b200", choices=["b200", "nccl"])
    ap.add_argument("--rounds", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-graph", action="store_true", help="eager local steps (for ncu launch lists)")
    ap.add_argument("--local-steps", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--seq", type=int, default=None)
    ap.add_argument("--server-mode", default="sharded")
    ap.add_argument("--out", default="")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    rank, world, lr_ = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr_)
    dev = torch.device("cuda", lr_)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from vantage6_b200.utils.timing import DeviceTimer, barrier_sync, max_over_ranks

    b200 = args.impl == "b200"
    res = {"model": args.model, "impl": args.impl, "world": world}
    torch.manual_seed(7)
    if args.model == "glm":
        from vantage6_b200.models.glm import FederatedGLM, synthetic_glm_shard

        rows = 1_000_000 // max(world, 1)
        X, y, w_true = synthetic_glm_shard(rows, 256, seed=100 + rank, device=dev)
        if b200:
            glm = FederatedGLM(X, y, rank, world, lr=2.0)
            step = glm.step
        else:       # baseline: two cuBLAS GEMVs + elementwise + ncclAllReduce of the 259-float payload
            w = torch.zeros(257, device=dev)
            Xf = X                                     # bf16 data, like the product arm
            last = {}

            def step():
                z = (Xf @ w[:256].to(Xf.dtype)).float() + w[256]
                r = torch.sigmoid(z) - y
                g = torch.cat([(Xf.t() @ r.to(Xf.dtype)).float(), r.sum()[None],
                               torch.nn.functional.binary_cross_entropy_with_logits(z, y, reduction="sum")[None],
                               torch.tensor([float(rows)], device=dev)])
                if world > 1:
                    dist.all_reduce(g)
                w.add_(g[:257] / g[258], alpha=-2.0)
                last["loss"] = g[257] / g[258]
                return last["loss"]
        for _ in range(20):
            step()
        barrier_sync(dev)
        t = DeviceTimer(dev)
        t.start()
        iters = 200
        for _ in range(iters):
            loss = step()
        ms = max_over_ranks(t.stop(), dev)
        res.update(us_per_iteration=1e3 * ms / iters, iterations_per_sec=iters / (ms / 1e3), loss=float(loss.item()),
                   rows_per_node=rows, bytes_X_per_iter=rows * 256 * 2,
                   X_read_GBps=rows * 256 * 2 / (ms / iters) / 1e6)
        if b200:
            res["coef_err"] = float((glm.w - w_true).abs().max().item())
    else:
        from vantage6_b200.models import zoo

        over = {}
        tr, spec = zoo.build_trainer(args.model, rank=rank, world=world, device=dev, server_mode=args.server_mode,
                                     data_plane="native" if b200 else "collective", fused_local_optimizer=b200,
                                     use_cuda_graph=(False if args.no_graph else None) if b200 else False, **over)
        n_steps = args.local_steps or spec.local_steps
        bsz = args.batch or spec.batch
        batches = [(x.to(dev), y.to(dev)) for x, y in spec.make_batches(n_steps, bsz, seed=500 + rank)]
        tr.initialize_global()
        for _ in range(args.warmup):
            tr.run_round(batches)
        barrier_sync(dev)
        t = DeviceTimer(dev)
        t.start()
        for _ in range(args.rounds):
            loss = tr.run_round(batches)
        ms = max_over_ranks(t.stop(), dev)
        if os.environ.get("V6_PROFILE_RANGE"):      # ncu --profile-from-start off: one more round, every thread's launches
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
            tr.run_round(batches)
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        # aggregation-only time (the communication-bound part)
        barrier_sync(dev)
        t.start()
        for _ in range(10):
            tr.engine.aggregate(1.0)
        agg_ms = max_over_ranks(t.stop(), dev) / 10
        res.update(ms_per_round=ms / args.rounds, rounds_per_sec=args.rounds / (ms / 1e3), node_rounds_per_sec=world * args.rounds / (ms / 1e3),
                   local_steps=n_steps, batch=bsz, loss=float(loss.item()), aggregate_ms=agg_ms,
                   n_federated_params=tr.fm.n_total, upload=tr.upload_mode, data_plane=tr.engine.data_plane,
                   multicast=bool(tr.engine.use_multicast), comm_status=tr.engine.poll_status())
        tr.close()
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(line + "\n")
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
