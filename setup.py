"""Packaging (parity: reference setup.py:29-64 -- version read from _version.py, ``__build__`` and
``rabbitmq/rabbitmq.config`` shipped as package data, console scripts ``vnode`` / ``vserver``;
plus the runtime entry points the reference's images provide: ``vnode-local`` / ``vserver-local``)."""
import os
from setuptools import find_packages, setup

here = os.path.abspath(os.path.dirname(__file__))
version_ns = {"__file__": os.path.join(here, "vantage6_b200", "_version.py")}
with open(version_ns["__file__"]) as f:
    exec(f.read(), version_ns)

setup(
    name="vantage6-b200",
    version=version_ns["__version__"],
    description="Blackwell-native federated-learning engine with vantage6's capabilities",
    packages=find_packages(include=["vantage6_b200", "vantage6_b200.*"]),
    python_requires=">=3.10",
    install_requires=["click", "pyyaml", "pyjwt", "cryptography", "numpy", "torch", "pyzmq"],
    extras_require={"events": ["websockets"], "tabular": ["pandas"], "test": ["pytest", "requests", "hypothesis", "scipy", "scikit-learn"]},
    package_data={"vantage6_b200": ["__build__", "cli/rabbitmq/rabbitmq.config", "ops/csrc/*", "ops/*.so"]},
    entry_points={"console_scripts": [
        "vnode=vantage6_b200.cli.node:cli_node",
        "vserver=vantage6_b200.cli.server:cli_server",
        "vnode-local=vantage6_b200.cli.node_local:main",
        "vserver-local=vantage6_b200.cli.server_local:main",
        "vdev=vantage6_b200.cli.dev:cli_dev",
    ]},
)
