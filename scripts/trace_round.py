"""Timeline of one federated round as it really runs (CUDA-graph replays, warm L2): kernel busy time, gaps between
kernels, time per kernel family.  Uses the CUPTI activity records of ``torch.profiler`` -- an ANALYSIS tool: durations
under a tracer are not bench values (the bench numbers come from bench.py), the point is where the round's wall time
goes that the per-kernel ncu list (serialised, cold caches) cannot show.

    PYTHONPATH=. python scripts/trace_round.py --model resnet50 --out gpurun_out/trace_round.json
"""
import argparse
import collections
import json
import re
import sys

sys.path.insert(0, ".")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--out", default="")
    ap.add_argument("--top", type=int, default=25)
    a = ap.parse_args()
    import torch
    from torch.profiler import ProfilerActivity, profile

    import bench

    args = bench.parse(["--model", a.model, "--steps", "1", "--warmup", "3", "--baselines", "", "--no-e2e"])
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    torch.manual_seed(1234)
    trainer, spec = bench.build_trainer("b200", args, 0, 1, dev)
    trainer.initialize_global()
    host, batches, B, n_steps, h2d, shape = bench.make_data(args, spec, 0, dev)
    n_samples = float(n_steps * B)
    for _ in range(3):
        trainer.run_round(batches, n_samples)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        trainer.run_round(batches, n_samples)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name) for e in evs if "memcpy" not in e.name.lower() and "memset" not in e.name.lower()),
                key=lambda t: t[0])
    if not ks:
        raise SystemExit("no kernel records")
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    busy, cur_s, cur_e = 0.0, ks[0][0], ks[0][1]
    gaps = []
    for s, e, n in ks[1:]:
        if s > cur_e:
            busy += cur_e - cur_s
            gaps.append((s - cur_e, n))
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    fam = collections.defaultdict(lambda: [0, 0.0])

    def family(name: str) -> str:
        m = re.match(r"(?:void )?((?:\w+::)*\w+(?:<[^>(]*>)?)", name)
        return (m.group(1) if m else name)[:70]

    for s, e, n in ks:
        f = fam[family(n)]
        f[0] += 1
        f[1] += e - s
    gap_by_next = collections.defaultdict(lambda: [0, 0.0])
    for g, n in gaps:
        f = gap_by_next[family(n)]
        f[0] += 1
        f[1] += g
    res = {"model": a.model, "kernels": len(ks), "span_us": t1 - t0, "busy_us": busy, "idle_us": (t1 - t0) - busy,
           "sum_of_kernel_us": sum(e - s for s, e, _ in ks),
           "families": sorted(([k, v[0], round(v[1], 1)] for k, v in fam.items()), key=lambda r: -r[2])[: a.top],
           "idle_before": sorted(([k, v[0], round(v[1], 1)] for k, v in gap_by_next.items()), key=lambda r: -r[2])[:12]}
    print(json.dumps(res, indent=1))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(res, f, indent=1)
    trainer.close()


if __name__ == "__main__":
    main()
