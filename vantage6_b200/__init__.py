"""vantage6_b200 -- a Blackwell-native federated-learning engine with vantage6's capabilities.

Layers (re-designed B200-first):
  cli/        vnode, vserver (same commands / options / messages as the reference CLI)
  common/     console printers, config manager + schema, AppContext, RSA encryption, prompts
  runtime/    process "containers": the Docker-daemon replacement (one process per GPU node)
  server/     central server: sqlite entity store, JWT auth, REST API, event channel
  node/       node runtime: task listener, algorithm runner, proxy server
  client/     UserClient / ContainerClient / ClientMockProtocol
  algorithm/  algorithm wrapper (master / RPC_ convention) + built-ins
  ops/        hand-written sm_100a kernels (K1-K8) + native symmetric heap
  parallel/   NVLink symmetric memory, FedAvg engine, federated trainer
  models/     ResNet-50, BERT-base, Llama-3 (+LoRA), logistic GLM on flat parameter buffers
  utils/      device timing, clocks, metrics, checkpoints
"""
from ._version import __version__, version_info  # noqa: F401
