"""K1 on the engine path (torchrun, >= 2 GPUs): FederatedTrainer(bcast="fused") -- the bf16 weights of the first-consumer
layers are left on their owners by the aggregation kernel and arrive inside the first forward GEMM of the round -- against
the plain path where K2 pushes everything.  Same seeds, same batches: the global models must agree after every round.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/dist_k1_engine_check.py
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("V6B200_LINEAR_BWD", "cublas")       # deterministic filter gradients: the two runs are compared closely
os.environ.setdefault("V6B200_LINEAR_FWD", "gemm")         # both runs on csrc/gemm.cu (the K1 kernel is that kernel): bit-comparable

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--model", default="bert_small")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", rank)))
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    from vantage6_b200.models import zoo
    from vantage6_b200.models.transformer import ShadowLinear

    def build(bcast):
        torch.manual_seed(0)
        spec = zoo.get(a.model)
        model = spec.build(dev)
        if world == 2:       # make the K1 prefix larger than one reducer slice: both ranks own blocks of some layer
            for name, m in model.named_modules():
                if isinstance(m, ShadowLinear) and name.endswith("ffn1"):
                    m.first_consumer = True
        from vantage6_b200.parallel.trainer import FederatedTrainer

        kw = dict(spec.trainer_kwargs)
        tr = FederatedTrainer(model, spec.forward_loss, rank=rank, world=world, device=dev, optimizer=spec.optimizer, lr=spec.lr,
                              upload=spec.upload, shadow_bf16=True, amp_dtype=torch.bfloat16, bcast=bcast, **kw)
        return tr, spec

    res = {"world": world, "model": a.model}
    trs = {}
    for mode in ("fused", "push"):
        tr, spec = build(mode)
        tr.initialize_global()
        trs[mode] = tr
    res["k1_layers"] = trs["fused"].k1_layers
    res["shadow_skip"] = list(trs["fused"].engine.shadow_skip)
    res["own_blocks"] = [list(m.k1["own_blocks"]) for m in trs["fused"].model.modules() if getattr(m, "k1", None)]
    assert trs["fused"].k1_layers > 0, "the K1 path did not engage"
    w0 = trs["push"].engine.w.clone()
    # same flat layout? no: the fused trainer orders the K1 weights first -- compare through the named views
    batches = [(x.to(dev), y.to(dev)) for x, y in spec.make_batches(spec.local_steps, spec.batch, 1000 + rank)]
    rows = []
    for r in range(a.rounds):
        losses = {m: float(trs[m].run_round(batches, float(spec.batch * spec.local_steps * (rank + 1))).item()) for m in ("fused", "push")}
        torch.cuda.synchronize()
        vf, vp = trs["fused"].fm.views(), trs["push"].fm.views()
        worst, moved = 0.0, 0.0
        for k in vp:
            d = (vf[k].float() - vp[k].float()).abs().max().item()
            worst = max(worst, d / (vp[k].float().abs().max().item() + 1e-6))
        sf, sp = trs["fused"].fm.shadow_views(), trs["push"].fm.shadow_views()
        shadow_worst, shadow_key = 0.0, ""
        for k in sp:
            d = (sf[k].float() - sp[k].float()).abs().max().item()
            if d > shadow_worst:
                shadow_worst, shadow_key = d, k
        rows.append({"round": r, "loss_fused": losses["fused"], "loss_push": losses["push"], "worst_rel_diff": worst, "shadow_abs_diff": shadow_worst, "shadow_key": shadow_key,
                     "status": trs["fused"].engine.poll_status()})
        assert trs["fused"].engine.poll_status() == 0
        assert abs(losses["fused"] - losses["push"]) < 2e-2 * max(1.0, abs(losses["push"])), rows[-1]
        assert worst < 2e-2, rows[-1]
    res["rounds"] = rows
    res["ok"] = True
    flag = torch.ones(1, device=dev)
    dist.all_reduce(flag)
    if rank == 0:
        line = json.dumps(res)
        print(line, flush=True)
        if a.out:
            os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
            with open(a.out, "w") as f:
                f.write(line + "\n")
    for tr in trs.values():
        tr.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
