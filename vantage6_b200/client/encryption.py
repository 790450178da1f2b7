"""``vantage6.client.encryption`` import path (reference vantage6/cli/node.py:44)."""
from ..common.encryption import CryptorBase, DummyCryptor, RSACryptor  # noqa: F401
