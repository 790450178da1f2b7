"""Checkpoint / resume (absent as a feature in the reference, whose only persistent
state is the server database and YAML configs).

A federated run persists, per rank: the global model slice owner state (fp32 master + server
optimizer moments + round counter), the local optimizer state and the node's parameter buffer, as
one ``safetensors`` file plus a small JSON header.  Resume = load + (the next aggregation
re-broadcasts the global model, so nodes that restart from an older file converge after one round).
"""
from __future__ import annotations

import json
from pathlib import Path
from typing import Any, Dict

import torch
from safetensors.torch import load_file, save_file


def save_checkpoint(path, trainer, round_idx: int, extra: Dict[str, Any] | None = None) -> Path:
    path = Path(path)
    path.mkdir(parents=True, exist_ok=True)
    eng = trainer.engine
    tensors = {"w": eng.w.detach().cpu().clone(), "w_global": eng.w_global.detach().cpu().clone(),
               "opt_m": eng.opt_m.detach().cpu().clone(), "opt_v": eng.opt_v.detach().cpu().clone()}
    local = trainer.opt.state_dict() if trainer.opt is not None else {}
    for k, v in local.items():
        if torch.is_tensor(v):
            tensors[f"local.{k}"] = v.detach().cpu().clone()
    if getattr(trainer, "w_ref", None) is not None:
        tensors["w_ref"] = trainer.w_ref.detach().cpu().clone()
    f = path / f"rank{eng.rank}.safetensors"
    save_file(tensors, str(f))
    if getattr(trainer, "torch_opt", None) is not None:       # stock-optimizer arm: its state is not a flat tensor set
        torch.save(trainer.torch_opt.state_dict(), str(path / f"rank{eng.rank}.torch_opt.pt"))
    meta = {"round": int(round_idx), "server_step": int(eng.server_step), "epoch": int(eng.epoch), "world": eng.world,
            "rank": eng.rank, "server_mode": eng.server_mode, "server_opt": eng.opt.name, "lo": eng.lo, "hi": eng.hi,
            "reducers": list(getattr(eng, "reducers", [])), "live_mask": int(eng.live_mask),
            "local_steps": int(local.get("steps", 0)) if local else 0, "extra": extra or {}}
    (path / f"rank{eng.rank}.json").write_text(json.dumps(meta))
    return path


def load_checkpoint(path, trainer) -> Dict[str, Any]:
    path = Path(path)
    eng = trainer.engine
    meta = json.loads((path / f"rank{eng.rank}.json").read_text())
    if meta["world"] != eng.world or meta["server_mode"] != eng.server_mode:
        raise ValueError(f"checkpoint was written for world={meta['world']} mode={meta['server_mode']}, "
                         f"this run is world={eng.world} mode={eng.server_mode}")
    t = load_file(str(path / f"rank{eng.rank}.safetensors"))
    with torch.no_grad():
        eng.w.copy_(t["w"])
        eng.w_global.copy_(t["w_global"])
        eng.opt_m.copy_(t["opt_m"])
        eng.opt_v.copy_(t["opt_v"])
        if eng.shadow is not None:
            eng.shadow.copy_(eng.w.to(torch.bfloat16))
        if "w_ref" in t and getattr(trainer, "w_ref", None) is not None:
            trainer.w_ref.copy_(t["w_ref"])
    eng.server_step = int(meta["server_step"])
    topt = path / f"rank{eng.rank}.torch_opt.pt"
    if getattr(trainer, "torch_opt", None) is not None and topt.exists():
        trainer.torch_opt.load_state_dict(torch.load(str(topt), map_location=eng.device))
    if trainer.opt is not None:
        sd = {k[len("local."):]: v for k, v in t.items() if k.startswith("local.")}
        sd["steps"] = meta["local_steps"]
        trainer.opt.load_state_dict(sd)
    return meta
