"""Regenerates docs/INVENTORY.md: where each item of SURVEY.md sections 2 and 5 lives, with the file:line of the defining
symbol looked up in the tree (fails loudly when a symbol has moved).

    python scripts/inventory.py > docs/INVENTORY.md
"""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def loc(path, pattern):
    """`path:line` of the first line of ``path`` that matches ``pattern``."""
    for i, line in enumerate((ROOT / path).read_text().splitlines(), 1):
        if re.search(pattern, line):
            return f"`{path}:{i}`"
    raise SystemExit(f"NOT FOUND {path} {pattern}")


rows=[
("C1","Packaging, console scripts",[("setup.py",r"entry_points")],"`vnode`, `vserver`, `vnode-local`, `vserver-local`; `requirements.txt`"),
("C2","Version / `__build__`",[("vantage6_b200/_version.py",r"^version_info =")],"`__version__` (PEP 440 from `version_info` + `__build__`)"),
("C3","Globals",[("vantage6_b200/cli/globals.py",r"DEFAULT_SERVER_SYSTEM_FOLDERS")],"same names and values as the reference"),
("C4","Utils (name rule, runtime ping)",[("vantage6_b200/cli/instance.py",r"def check_config_name_allowed"),("vantage6_b200/cli/instance.py",r"def check_if_docker_deamon_is_running")],"re-exported by `cli/utils.py`"),
("C5","Config schema + managers",[("vantage6_b200/cli/configuration_manager.py",r"class ServerConfiguration"),("vantage6_b200/cli/configuration_manager.py",r"class NodeConfiguration"),("vantage6_b200/common/schema.py",r"^class Schema:"),("vantage6_b200/common/configuration_manager.py",r"^class ConfigurationManager")],"own `schema` implementation (the package is not installed)"),
("C6","ServerContext",[("vantage6_b200/cli/context.py",r"^class ServerContext")],"`get_database_uri`, `docker_container_name`, env overrides"),
("C7","NodeContext",[("vantage6_b200/cli/context.py",r"^class NodeContext")],"volume / network / container names, `databases`"),
("C8","Node wizard",[("vantage6_b200/cli/configuration_wizard.py",r"def node_configuration_questionaire")],"prompts via `common/prompts.py` (questionary is not installed)"),
("C9","Server wizard",[("vantage6_b200/cli/configuration_wizard.py",r"def server_configuration_questionaire")],""),
("C10","Wizard driver",[("vantage6_b200/cli/configuration_wizard.py",r"^def configuration_wizard")],"merges an environment into an existing file"),
("C11","Config selector",[("vantage6_b200/cli/configuration_wizard.py",r"def select_configuration_questionaire")],""),
("C12","`vnode list`",[("vantage6_b200/cli/node.py",r"def cli_node_list")],"byte-exact table (`tests/test_node_cli.py`)"),
("C13","`vnode new`",[("vantage6_b200/cli/node.py",r"def cli_node_new_configuration")],""),
("C14","`vnode files`",[("vantage6_b200/cli/node.py",r"def cli_node_files")],""),
("C15","`vnode start`",[("vantage6_b200/cli/node.py",r"def cli_node_start")],"+ `--gpu K`; process runtime `runtime/__init__.py` in place of the docker SDK"),
("C16","`vnode stop`",[("vantage6_b200/cli/node.py",r"def cli_node_stop")],""),
("C17","`vnode attach`",[("vantage6_b200/cli/node.py",r"def cli_node_attach")],""),
("C18","`vnode create-private-key`",[("vantage6_b200/cli/node.py",r"def cli_node_create_private_key"),("vantage6_b200/common/encryption.py",r"^class RSACryptor")],"`PATCH /organization/<id>` with the public key"),
("C19","`vnode clean`",[("vantage6_b200/cli/node.py",r"def cli_node_clean")],""),
("C20","`vnode remove`",[("vantage6_b200/cli/node.py",r"def cli_node_remove")],""),
("C21","`vnode version`",[("vantage6_b200/cli/node.py",r"def cli_node_version")],""),
("C22","`click_insert_context`",[("vantage6_b200/cli/server.py",r"def click_insert_context")],""),
("C23","`vserver start`",[("vantage6_b200/cli/server.py",r"def cli_server_start")],"starts `vserver-local start` + the message-queue sidecar"),
("C24","`vserver list`",[("vantage6_b200/cli/server.py",r"def cli_server_configuration_list")],""),
("C25","`vserver files`",[("vantage6_b200/cli/server.py",r"def cli_server_files")],""),
("C26","`vserver new`",[("vantage6_b200/cli/server.py",r"def cli_server_new")],""),
("C27","`vserver import`",[("vantage6_b200/cli/server.py",r"def cli_server_import"),("vantage6_b200/server/fixtures.py",r"^def load")],""),
("C28","`vserver shell / stop / attach / version`",[("vantage6_b200/cli/server.py",r"def cli_server_shell"),("vantage6_b200/cli/server.py",r"def cli_server_stop"),("vantage6_b200/cli/server.py",r"def cli_server_attach"),("vantage6_b200/cli/server.py",r"def cli_server_version")],"the reference's `-server` suffix bug in `version` is not reproduced"),
("C29","RabbitMQ manager",[("vantage6_b200/cli/rabbitmq/queue_manager.py",r"^class RabbitMQManager"),("vantage6_b200/cli/rabbitmq/queue_manager.py",r"def split_rabbitmq_uri"),("vantage6_b200/server/mq_broker.py",r"^def attach_app")],"drives an in-tree ZeroMQ XSUB/XPUB broker; events mirrored between server processes"),
("C30","Test runner",[("utest.py",r"def run"),("vantage6_b200/common/utest.py",r"def find_tests")],""),
("C31","Tests",[],"`tests/test_node_cli.py`, `test_server_cli.py`, `test_wizard.py`, `test_config_context.py`, `test_cli_instance.py` + runtime / server / plumbing / engine / GPU tiers"),
("C32","Build / CI",[],"`Makefile`, `.github/workflows/ci.yaml`, `requirements.txt`, `vantage6_b200/ops/build.py`"),
]
print("# Component inventory: SURVEY.md section 2, line by line\n")
print("Where every component of the reference (SURVEY.md 2.1), every external symbol its call sites rely on (2.2), every plane of\nits communication backend (2.5), every kernel of the work-list (2.6) and every auxiliary subsystem (5) lives in this repository.\nGenerated from the tree (`file:line` of the defining symbol), so the references hold for this commit.\n")
print("## 2.1 Components of the reference\n\n| # | component | here | notes |\n|---|---|---|---|")
for cid,name,locs,note in rows:
    print(f"| {cid} | {name} | {', '.join(loc(*l) for l in locs) if locs else '-'} | {note} |")

ext=[
("`vantage6.common.{info,warning,error,debug}`",[("vantage6_b200/common/__init__.py",r"^def info")],"`[info]  - msg` format asserted by the CLI tests"),
("`bytes_to_base64s`, `check_config_write_permissions`, `STRING_ENCODING`",[("vantage6_b200/common/__init__.py",r"def bytes_to_base64s"),("vantage6_b200/common/__init__.py",r"def check_config_write_permissions")],""),
("`globals.APPNAME`, `DEFAULT_DOCKER_REGISTRY`, `DEFAULT_NODE_IMAGE`, `DEFAULT_SERVER_IMAGE`, `VPN_CONFIG_FILE`",[("vantage6_b200/common/globals.py",r"^APPNAME")],""),
("`AppContext`",[("vantage6_b200/common/context.py",r"^class AppContext")],"folders per scope, rotating log files, `LOGGING_ENABLED`"),
("`Configuration`, `ConfigurationManager`",[("vantage6_b200/common/configuration_manager.py",r"^class Configuration\b"),("vantage6_b200/common/configuration_manager.py",r"^class ConfigurationManager")],"one YAML, several environments"),
("`docker.addons` (`pull_if_newer`, `remove_container_if_exists`, `check_docker_running`, `get_server_config_name`)",[("vantage6_b200/runtime/addons.py",r"def pull_if_newer")],"over the process runtime"),
("`NetworkManager`",[("vantage6_b200/runtime/addons.py",r"class NetworkManager")],""),
("`utest.{find_tests, run_tests}`",[("vantage6_b200/common/utest.py",r"def run_tests")],""),
("`vantage6.client.Client`",[("vantage6_b200/client/__init__.py",r"^class UserClient"),("vantage6_b200/client/__init__.py",r"^class ContainerClient"),("vantage6_b200/client/mock.py",r"^class ClientMockProtocol")],"endpoint / call table: `docs/SERVER_API.md`"),
("`RSACryptor`",[("vantage6_b200/common/encryption.py",r"^class RSACryptor")],"vantage6 wire format `key$iv$ciphertext`: AES-256-CTR payload, its key sealed with the receiving organization's RSA key (PKCS#1 v1.5)"),
("server runtime (`vserver-local`, WSGI `app`)",[("vantage6_b200/cli/server_local.py",r"def cli_server_local"),("vantage6_b200/server/app.py",r"^class ServerApp"),("vantage6_b200/server/admin_routes.py",r"^def register"),("vantage6_b200/server/ws_events.py",r"^class WebSocketEvents"),("vantage6_b200/server/db.py",r"^class Database")],"REST + JWT + rules, websocket events, sqlite"),
("node runtime (`vnode-local`)",[("vantage6_b200/cli/node_local.py",r"def cli_node_local"),("vantage6_b200/node/__init__.py",r"^class Node\b"),("vantage6_b200/node/proxy.py",r"^class ProxyServer"),("vantage6_b200/node/zygote.py",r"^class Zygote:"),("vantage6_b200/node/gpu_worker.py",r"^class GpuWorker")],"runs algorithms as child processes with the container env / file contract; resident GPU worker"),
("algorithm interface",[("vantage6_b200/algorithm/wrapper.py",r"^def dispatch"),("vantage6_b200/algorithm/__init__.py",r"^IMAGES"),("vantage6_b200/algorithm/data.py",r"^def make_local_batches"),("vantage6_b200/algorithm/peer.py",r"^class PeerChannel")],"`master` / `RPC_`, data loaders, node-to-node channel; built-ins in `algorithm/builtin/`"),
]
print("\n## 2.2 External-package contract\n\n| symbol(s) | here | notes |\n|---|---|---|")
for name,locs,note in ext:
    print(f"| {name} | {', '.join(loc(*l) for l in locs)} | {note} |")

planes=[
("control plane user -> server",[("vantage6_b200/client/__init__.py",r"^class ClientBase"),("vantage6_b200/common/jsonhttp.py",r"^class JsonHttp")],"REST + JWT over keep-alive connections"),
("control plane server <-> node",[("vantage6_b200/server/ws_events.py",r"^class WebSocketEvents"),("vantage6_b200/node/__init__.py",r"def _listen_websocket")],"websocket push, long-poll fallback; no tensor bytes"),
("server <-> server fan-out",[("vantage6_b200/server/mq_broker.py",r"^def attach_app")],"ZeroMQ broker sidecar behind the `rabbitmq_uri` key"),
("data plane (model broadcast / delta upload)",[("vantage6_b200/parallel/symm.py",r"^class SymmetricHeap"),("vantage6_b200/ops/csrc/symm.cpp",r"p_cuMemCreate\("),("vantage6_b200/parallel/fedavg.py",r"^class FedAvgEngine")],"CUDA VMM symmetric heap, fd passing, NVLS multicast objects, signal pads"),
("node <-> node algorithm traffic",[("vantage6_b200/algorithm/peer.py",r"^class PeerChannel")],"symmetric memory + the small all-reduce kernel; gloo on CPU nodes"),
("collectives baseline",[("bench.py",r"^def run_trainer_arm")],"NCCL arm and stock-graph arm run by `bench.py` after the product arm"),
]
print("\n## 2.5 Communication planes\n\n| plane | here | how |\n|---|---|---|")
for name,locs,note in planes:
    print(f"| {name} | {', '.join(loc(*l) for l in locs)} | {note} |")

kern=[
("K1","broadcast fused with the first consuming GEMM",[("vantage6_b200/ops/csrc/gemm.cu",r"k1_push"),("vantage6_b200/parallel/fedavg.py",r"def k1_layer"),("vantage6_b200/models/transformer.py",r"^K1_STEP")],"`FederatedTrainer(bcast=\"fused\")`; `tests/dist_k1_engine_check.py`"),
("K2","weighted reduction + server optimizer + broadcast",[("vantage6_b200/ops/csrc/fedavg.cu",r"fedavg_round_kernel")],"`multimem.ld_reduce` / P2P, FedAvg / FedAvgM / FedAdam in registers"),
("K3","small-message aggregation",[("vantage6_b200/ops/csrc/fedavg.cu",r"small_allreduce_kernel")],"one CTA, signal pads"),
("K4","attention forward / backward",[("vantage6_b200/ops/csrc/attention2.cu",r"flash_fwd2_kernel"),("vantage6_b200/ops/csrc/attention_bwd.cu",r"flash_bwd_dkv_kernel")],"tcgen05 + TMEM + TMA, V read in place"),
("K5","LayerNorm / RMSNorm",[("vantage6_b200/ops/csrc/norm.cu",r"norm_fwd_kernel")],"split backward (row-wise dx, column-wise parameter gradients)"),
("K6","RoPE",[("vantage6_b200/ops/csrc/rope_glm.cu",r"rope")],""),
("K7","flat SGD / AdamW + delta publish",[("vantage6_b200/ops/csrc/optim.cu",r"flat_optim_kernel")],"one launch per step over the flat parameter buffer"),
("K8","logistic GLM step",[("vantage6_b200/ops/csrc/glm_tc.cu",r"glm_tc_kernel")],"two-launch iteration with the NVLink all-reduce"),
("X1","convolutions of ResNet-50 (fprop / dgrad / wgrad, stem, stride 2)",[("vantage6_b200/ops/csrc/igemm.cu",r"igemm_kernel")],"TMA im2col implicit GEMM, BN statistics epilogue, split-K wgrad"),
("X2","transformer backward GEMMs, tied head, cross-entropy",[("vantage6_b200/ops/conv.py",r"def linear_dgrad"),("vantage6_b200/ops/csrc/ce.cu",r"ce_fwd_kernel")],"no cuBLAS launch left in a BERT round"),
("BN","BatchNorm apply / backward, pooling",[("vantage6_b200/ops/csrc/bn.cu",r"bn_apply_kernel"),("vantage6_b200/ops/csrc/pool.cu",r"maxpool_bwd2x2_kernel")],""),
]
print("\n## 2.6 Kernel work-list (details, tests and measurements: `docs/KERNELS.md`)\n\n| id | kernel | here | notes |\n|---|---|---|---|")
for kid,name,locs,note in kern:
    print(f"| {kid} | {name} | {', '.join(loc(*l) for l in locs)} | {note} |")

aux=[
("5.1 tracing / profiling",[("vantage6_b200/utils/timing.py",r"^class DeviceTimer"),("scripts/trace_round.py",r"def main"),("scripts/ncu_summary.py",r"^def load"),("scripts/sass_report.py",r"^def main|^def report|^def ")],"CUDA-event timers, NVTX ranges in `bench.py`, round timeline, ncu / SASS / ptxas summaries; `V6B200_TRACE_TASKS`, `V6B200_TRACE_HTTP`, `GET /metrics`"),
("5.2 sanitizers",[("scripts/sanitize.sh",r"timeout 600 compute-sanitizer")],"logs: `profiles/sanitizer_*_r2.txt`"),
("5.3 failure detection / recovery",[("vantage6_b200/parallel/fedavg.py",r"def mark_dead"),("vantage6_b200/parallel/trainer.py",r"def recover_if_failed"),("vantage6_b200/server/app.py",r"def reap_silent_nodes"),("tests/dist_fault_check.py",r"^def main")],"bounded spins + abort word in the kernels, kill-a-rank test at 2 and 8 GPUs, heartbeats + reaper"),
("5.4 checkpoint / resume",[("vantage6_b200/utils/checkpoint.py",r"^def save_checkpoint"),("vantage6_b200/algorithm/builtin/fedavg.py",r"checkpoint_every")],"`checkpoint_every` / `resume_from` task kwargs; server database survives restarts"),
("5.5 metrics / logging",[("vantage6_b200/utils/metrics.py",r"^class MetricsWriter"),("vantage6_b200/common/context.py",r"RotatingFileHandler"),("vantage6_b200/server/admin_routes.py",r"def metrics")],"per-round JSONL in the node log directory, rotating instance logs, Prometheus endpoint"),
("5.6 config / flags",[("vantage6_b200/common/configuration_manager.py",r"^class ConfigurationManager"),("vantage6_b200/cli/context.py",r"os.environ.get\(\"VANTAGE6_DB_URI")],"multi-environment YAML, scopes, env overrides; environment switches listed in `docs/ALGORITHMS.md`"),
("5.8 communication backend",[("vantage6_b200/ops/csrc/symm.cpp",r"p_cuMulticastCreate\("),("vantage6_b200/parallel/symm.py",r"^class SymmetricHeap")],""),
]
print("\n## 5 Auxiliary subsystems\n\n| subsystem | here | notes |\n|---|---|---|")
for name,locs,note in aux:
    print(f"| {name} | {', '.join(loc(*l) for l in locs)} | {note} |")
