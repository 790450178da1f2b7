"""``ClientMockProtocol`` -- test an algorithm locally without a server or nodes: the ``master``
gets this client, ``create_new_task`` runs the ``RPC_`` partials in-process over a list of
datasets (one per simulated organization) and ``get_results`` returns their outputs.  This is the
upstream answer to "test multi-node without a cluster" (SURVEY.md Appendix C / section 4).
"""
from __future__ import annotations

import importlib
from typing import Any, Dict, List, Sequence


class ClientMockProtocol:
    def __init__(self, datasets: Sequence[Any], module: str | Any):
        self.n = len(datasets)
        self.datasets = list(datasets)
        self.lib = importlib.import_module(module) if isinstance(module, str) else module
        self.tasks: List[Dict[str, Any]] = []
        self.organization_id = 0
        self.collaboration_id = 0
        self.host_node_id = 0
        self.image = getattr(self.lib, "__name__", "mock")

    def create_new_task(self, input_: dict, organization_ids: Sequence[int] = (), name: str = "mock",
                        description: str | None = None, **_) -> dict:
        if input_.get("master"):
            result = [getattr(self.lib, input_["method"])(self, self.datasets[0], *input_.get("args", []),
                                                         **input_.get("kwargs", {}))]
        else:
            method = getattr(self.lib, f"RPC_{input_['method']}")
            ids = list(organization_ids) or list(range(self.n))
            result = [method(self.datasets[i], *input_.get("args", []), **input_.get("kwargs", {})) for i in ids]
        task = {"id": len(self.tasks), "results": result, "complete": True}
        self.tasks.append(task)
        return task

    def get_task(self, task_id: int) -> dict:
        return {"id": task_id, "complete": True}

    def wait_for_task(self, task_id: int, **_) -> dict:
        return self.get_task(task_id)

    def get_results(self, task_id: int) -> List[Any]:
        return self.tasks[task_id]["results"]

    def wait_for_results(self, task_id: int, **_) -> List[Any]:
        return self.get_results(task_id)

    def get_organizations_in_my_collaboration(self) -> List[dict]:
        return [{"id": i, "name": f"mock-{i}"} for i in range(self.n)]

    def get_algorithm_addresses(self, task_id: int) -> List[dict]:
        return [{"rank": i, "node_id": i, "organization_id": i, "gpu": None} for i in range(self.n)]
