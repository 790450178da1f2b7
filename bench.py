#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): federated rounds/sec for ResNet-50 FedAvg, one federated
node per GPU, 1 local epoch per round on synthetic ImageNet-shape data, random-init weights.

    python bench.py --gpus N --steps K --warmup W            # this framework (fused kernels)
    python bench.py --impl nccl ...                          # in-repo NCCL + torch.optim baseline
    python bench.py --impl reference ...                     # the unmodified reference (unavailable)

A bench "step" is ONE federated round = `local_steps` local SGD steps on every node + the
server aggregation (weighted FedAvg reduce + server optimizer + broadcast of the new global).
`value` = nodes x rounds/sec (whole-job aggregate: local epochs completed per second across
the federation; `rounds_per_sec` is reported next to it).  Weak scaling: per-node work fixed.

Timing: W untimed warm-up rounds, then exactly K rounds between CUDA events, bracketed by
barrier + torch.cuda.synchronize() on both sides, max over ranks.  Inputs per round (77 MB
uint8 images + 102 MB fp32 model + GBs of activations) exceed the 126 MB L2, so no flush is
needed between rounds.  `e2e` repeats the measurement through the public API with the
round's batches living in pinned host memory (H2D copy every local step) and a device->host
read of the round's loss.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8, help="timed federated rounds")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "nccl", "reference"])
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "resnet_tiny"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--local-samples", type=int, default=512, help="node-local dataset size = 1 epoch")
    ap.add_argument("--server-mode", default="sharded", choices=["sharded", "central"])
    ap.add_argument("--server-opt", default="fedavg", choices=["fedavg", "fedavgm", "fedadam"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    return ap.parse_args()


def reference_unavailable():
    print(json.dumps({
        "impl": "reference",
        "unavailable": "reference snapshot is only the vantage6 CLI: --no-deps install imports fail "
                       "(questionary, docker, schema, vantage6.common, vantage6.client missing offline); it has no "
                       "FedAvg/NCCL/GPU path to time and needs a Docker daemon (see DESIGN.md)"}))


def main():
    args = parse()
    if args.impl == "reference":
        reference_unavailable()
        return 0

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and rank == 0:
        print(f"[bench] note: WORLD_SIZE={world} overrides --gpus {args.gpus}", file=sys.stderr)
    if not torch.cuda.is_available():
        print(json.dumps({"metric": "federated_node_rounds_per_sec", "value": None, "unit": "node-rounds/s",
                          "error": "no CUDA device visible"}))
        return 1
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from vantage6_b200.models.resnet import imagenet_forward_loss, resnet50, resnet_tiny
    from vantage6_b200.parallel.fedavg import ServerOptConfig
    from vantage6_b200.parallel.trainer import FederatedTrainer
    from vantage6_b200.utils.timing import ClockSampler, DeviceTimer, barrier_sync, max_over_ranks

    torch.manual_seed(1234)           # identical random init everywhere (rank 0's is authoritative)
    torch.backends.cudnn.benchmark = True
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    b200 = args.impl == "b200"
    # product arm: hand-written fused BN(+add)(+ReLU) kernels; baseline arm: stock BatchNorm2d / ReLU / add
    model = (resnet50(fused_bn=b200) if args.model == "resnet50" else resnet_tiny(1000, fused_bn=b200))
    model = model.to(memory_format=torch.channels_last)
    trainer = FederatedTrainer(
        model, imagenet_forward_loss, rank=rank, world=world, device=device, optimizer="sgd", lr=0.05, momentum=0.9,
        weight_decay=1e-4, server_mode=args.server_mode, server_opt=ServerOptConfig(args.server_opt, 1.0),
        upload="weights_f32", data_plane="native" if b200 else "collective",
        use_cuda_graph=b200 and not args.no_graph, fused_local_optimizer=b200, amp_dtype=torch.bfloat16,
        shadow_bf16=b200)       # product arm: conv filters are consumed from the bf16 shadow kept by K7 / K2
    trainer.initialize_global()

    B = args.batch
    n_steps = max(1, args.local_samples // B)
    res = 224 if args.model == "resnet50" else 64
    g = torch.Generator().manual_seed(100 + rank)       # every node has its own (non-IID) synthetic shard
    host_x = torch.randint(0, 256, (n_steps, B, 3, res, res), dtype=torch.uint8, generator=g).pin_memory()
    host_y = torch.randint(0, 1000, (n_steps, B), dtype=torch.int64, generator=g).pin_memory()
    dev_x, dev_y = host_x.to(device), host_y.to(device)
    dev_batches = [(dev_x[i], dev_y[i]) for i in range(n_steps)]
    host_batches = [(host_x[i], host_y[i]) for i in range(n_steps)]
    n_samples = float(n_steps * B)

    def run(batches, rounds, read_loss):
        last = None
        for _ in range(rounds):
            last = trainer.run_round(batches, n_samples)
            if read_loss:
                last = last.item()          # device -> host read of the round's result
        return last

    # ---------------- device-timed headline ----------------
    run(dev_batches, args.warmup, False)
    barrier_sync(device)
    timer = DeviceTimer(device)
    with ClockSampler(local_rank) as clocks:
        barrier_sync(device)
        launches0 = trainer.native_launches
        timer.start()
        torch.cuda.nvtx.range_push("v6_timed")          # ncu --nvtx --nvtx-include "v6_timed/"
        loss = run(dev_batches, args.steps, False)
        launches = trainer.native_launches - launches0
        torch.cuda.nvtx.range_pop()
        ms = timer.stop()
        barrier_sync(device)
    ms = max_over_ranks(ms, device)
    loss_val = float(loss.item())
    if os.environ.get("V6_PROFILE_RANGE"):      # ncu --profile-from-start off: one more round, all threads' launches
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        run(dev_batches, 1, False)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    status = trainer.engine.poll_status()

    # ---------------- end-to-end through the public API ----------------
    e2e = None
    if not args.no_e2e:
        run(host_batches, 1, True)
        barrier_sync(device)
        t0 = time.perf_counter()
        run(host_batches, args.steps, True)
        barrier_sync(device)
        e2e_s = max_over_ranks(time.perf_counter() - t0, device)
        e2e = {"value": world * args.steps / e2e_s, "unit": "node-rounds/s",
               "h2d_bytes_per_step": int(host_x[0].numel() * n_steps + host_y[0].numel() * 8 * n_steps),
               "d2h_bytes_per_step": 4, "ms_per_step": 1e3 * e2e_s / args.steps}

    rounds_per_sec = args.steps / (ms / 1e3)
    nv_bytes = trainer.engine.nvlink_bytes_per_round()
    if rank == 0:
        out = {
            "metric": "federated_node_rounds_per_sec (ResNet-50 FedAvg, rounds/sec x nodes)",
            "value": world * rounds_per_sec, "unit": "node-rounds/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": args.impl,
            "rounds_per_sec": rounds_per_sec, "images_per_sec": world * rounds_per_sec * n_samples,
            "config": {"model": args.model, "global_batch": B * world, "seq_len": None,
                       "parallelism": f"fedavg{world} (1 node/GPU, server {args.server_mode})",
                       "local_steps_per_round": n_steps, "local_batch": B, "local_samples": int(n_samples),
                       "image": [3, res, res], "server_opt": args.server_opt, "param_dtype": "fp32 master",
                       "l2": "inputs larger than L2 (77 MB images + 102 MB model per round), no flush",
                       "data_plane": trainer.engine.data_plane, "multicast": bool(trainer.engine.use_multicast),
                       "cuda_graph": bool(trainer.use_graph)},
            "clocks": clocks.summary(), "e2e": e2e,
            "gpu_launches": int(launches),
            "final_loss": loss_val, "comm_status": status,
            "nvlink_bytes_per_round_per_gpu": nv_bytes,
        }
        print(json.dumps(out), flush=True)
    trainer.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
