"""Algorithm wrapper -- the "algorithm-container interface" (SURVEY.md Appendix C).

An algorithm is a python module.  Its central part is ``master(client, data, *args, **kwargs)``;
its node-local parts are functions prefixed ``RPC_`` that receive the node's data.  The wrapper
dispatches on the task input ``{"method": name, "master": bool, "args": [...], "kwargs": {...}}``.

The node hands an algorithm run the same environment contract vantage6 gives a container
(the reference shows the node side of it: ``{LABEL}_DATABASE_URI`` / ``DATABASE_URI`` env vars
at reference vantage6/cli/node.py:352-378, per-run temp volumes at vantage6/cli/context.py:140-141,
the local proxy host ``proxyserver`` at vantage6/cli/globals.py:27):

    INPUT_FILE  OUTPUT_FILE  TOKEN_FILE  TEMPORARY_FOLDER  DATABASE_URI
    HOST  PORT  API_PATH            (address of the node's proxy server)
    V6_GPU                          (B200 extension: index of the GPU this node is pinned to)

Data loading understands ``.csv`` / ``.parquet`` (pandas), ``.pt`` (torch), ``.npy`` / ``.npz``
(numpy) and ``synthetic://...`` URIs (algorithms generate data on their own GPU).

Run as:  python -m vantage6_b200.algorithm.wrapper <module>
"""
from __future__ import annotations

import importlib
import logging
import os
import sys
import traceback
from typing import Any, Callable, Optional

from ..common.serialization import deserialize, serialize

log = logging.getLogger("wrapper")


def load_data(uri: Optional[str]) -> Any:
    """Load the node's database for an algorithm run."""
    if not uri:
        return None
    if "://" in uri and not uri.startswith("file://"):
        return uri                      # e.g. synthetic://imagenet?n=512 -- algorithm interprets it
    path = uri[7:] if uri.startswith("file://") else uri
    ext = os.path.splitext(path)[1].lower()
    if ext == ".csv":
        import pandas as pd

        return pd.read_csv(path)
    if ext == ".parquet":
        import pandas as pd

        return pd.read_parquet(path)
    if ext in (".pt", ".pth"):
        import torch

        return torch.load(path, map_location="cpu", weights_only=False)
    if ext == ".npy":
        import numpy as np

        return np.load(path)
    if ext == ".npz":
        import numpy as np

        return dict(np.load(path))
    if ext == ".json":
        with open(path, "rb") as f:
            return deserialize(f.read(), "json")
    raise ValueError(f"do not know how to load database {uri!r}")


def dispatch(module, input_data: dict, data: Any, client_factory: Callable[[], Any]) -> Any:
    """Call ``master`` (with a client) or ``RPC_<method>`` (with the data)."""
    method_name = input_data["method"]
    args = input_data.get("args", []) or []
    kwargs = input_data.get("kwargs", {}) or {}
    if input_data.get("master"):
        log.info("Running a master-container: %s", method_name)
        method = getattr(module, method_name)
        return method(client_factory(), data, *args, **kwargs)
    log.info("Running a regular container: RPC_%s", method_name)
    method = getattr(module, f"RPC_{method_name}", None)
    if method is None:
        raise AttributeError(f"method 'RPC_{method_name}' not found in {module.__name__}")
    return method(data, *args, **kwargs)


def run_algorithm(module_name: str) -> int:
    """Container entry point: read the env contract, dispatch, write OUTPUT_FILE."""
    logging.basicConfig(level=logging.INFO, stream=sys.stdout,
                        format="%(asctime)s - %(name)-14s - %(levelname)-8s - %(message)s")
    module = importlib.import_module(module_name)
    with open(os.environ["INPUT_FILE"], "rb") as f:
        input_data = deserialize(f.read())
    fmt = input_data.get("output_format", "json") if isinstance(input_data, dict) else "json"
    label = os.environ.get("DATABASE_LABEL", "default").upper()
    uri = os.environ.get(f"{label}_DATABASE_URI") or os.environ.get("DATABASE_URI")
    data = load_data(uri)

    def client_factory():
        from ..client import ContainerClient

        with open(os.environ["TOKEN_FILE"]) as f:
            token = f.read().strip()
        port = os.environ.get("PORT")
        return ContainerClient(token=token, host=os.environ["HOST"], port=int(port) if port else None,
                               path=os.environ.get("API_PATH", ""))

    try:
        output = dispatch(module, input_data, data, client_factory)
    except Exception:  # noqa: BLE001
        traceback.print_exc()
        return 1
    with open(os.environ["OUTPUT_FILE"], "wb") as f:
        f.write(serialize(output, fmt))
    return 0


# vantage6 3.x name of the entry point used by algorithm images
docker_wrapper = run_algorithm

if __name__ == "__main__":
    sys.exit(run_algorithm(sys.argv[1]))
