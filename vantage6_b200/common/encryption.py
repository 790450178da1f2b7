"""End-to-end encryption of task inputs / results between organizations.

Contract seen from the reference CLI (reference vantage6/cli/node.py:44,570-591):
``RSACryptor(file).private_key``, ``RSACryptor.create_new_rsa_key(path)``,
``RSACryptor.create_public_key_bytes(private_key)``; the public key travels base64-encoded in
``PATCH /organization/<id>`` (reference node.py:610-614).

Scheme (hybrid, like upstream vantage6): a random 256-bit AES key encrypts the payload
(AES-CTR), the AES key is encrypted with the receiving organization's RSA public key (PKCS1v15);
the message is ``b64(enc_key)$b64(iv)$b64(ciphertext)``.  ``DummyCryptor`` is used when a
collaboration is not encrypted: it only base64-encodes.

Note for the B200 data plane: tensors that move over NVLink symmetric memory never pass through
this path; it protects the *control-plane* payloads (task inputs, small results) exactly as in
vantage6.
"""
from __future__ import annotations

import logging
import os
from pathlib import Path

from cryptography.hazmat.backends import default_backend
from cryptography.hazmat.primitives import serialization
from cryptography.hazmat.primitives.asymmetric import padding, rsa
from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
from cryptography.hazmat.primitives.serialization import load_pem_private_key, load_pem_public_key

from . import base64s_to_bytes, bytes_to_base64s, logger_name

SEPARATOR = "$"


class CryptorBase:
    def __init__(self):
        self.log = logging.getLogger(logger_name(__name__))

    @staticmethod
    def bytes_to_str(data: bytes) -> str:
        return bytes_to_base64s(data)

    @staticmethod
    def str_to_bytes(data: str) -> bytes:
        return base64s_to_bytes(data)

    def encrypt_bytes_to_str(self, data: bytes, pubkey_base64: str) -> str:
        return self.bytes_to_str(data)

    def decrypt_str_to_bytes(self, data: str) -> bytes:
        return self.str_to_bytes(data)


class DummyCryptor(CryptorBase):
    """Does no encryption: only base64 (un-encrypted collaborations)."""


class RSACryptor(CryptorBase):
    def __init__(self, private_key_file):
        super().__init__()
        self.private_key = self.__load_private_key(private_key_file)

    def __load_private_key(self, private_key_file):
        p = Path(private_key_file)
        if not p.exists():
            raise FileNotFoundError(f"Private key file {private_key_file} not found.")
        self.log.debug("Loading private key")
        return load_pem_private_key(p.read_bytes(), password=None, backend=default_backend())

    # -- key management -------------------------------------------------------------------
    @staticmethod
    def create_new_rsa_key(path, bits: int = 4096):
        private_key = rsa.generate_private_key(backend=default_backend(), key_size=bits, public_exponent=65537)
        pem = private_key.private_bytes(encoding=serialization.Encoding.PEM,
                                        format=serialization.PrivateFormat.TraditionalOpenSSL,
                                        encryption_algorithm=serialization.NoEncryption())
        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        path.write_bytes(pem)
        try:
            os.chmod(path, 0o600)
        except OSError:
            pass
        return private_key

    @staticmethod
    def create_public_key_bytes(private_key) -> bytes:
        return private_key.public_key().public_bytes(encoding=serialization.Encoding.PEM,
                                                     format=serialization.PublicFormat.SubjectPublicKeyInfo)

    @property
    def public_key_bytes(self) -> bytes:
        return self.create_public_key_bytes(self.private_key)

    @property
    def public_key_str(self) -> str:
        return bytes_to_base64s(self.public_key_bytes)

    def verify_public_key(self, pubkey_base64: str) -> bool:
        """True if ``pubkey_base64`` (as stored at the server) matches our private key."""
        try:
            return base64s_to_bytes(pubkey_base64) == self.public_key_bytes
        except Exception:  # noqa: BLE001
            return False

    # -- hybrid encryption ----------------------------------------------------------------
    def encrypt_bytes_to_str(self, data: bytes, pubkey_base64s: str) -> str:
        pubkey = load_pem_public_key(base64s_to_bytes(pubkey_base64s), backend=default_backend())
        shared_key = os.urandom(32)
        iv = os.urandom(16)
        enc = Cipher(algorithms.AES(shared_key), modes.CTR(iv), backend=default_backend()).encryptor()
        ciphertext = enc.update(data) + enc.finalize()
        enc_key = pubkey.encrypt(shared_key, padding.PKCS1v15())
        return SEPARATOR.join([self.bytes_to_str(enc_key), self.bytes_to_str(iv), self.bytes_to_str(ciphertext)])

    def decrypt_str_to_bytes(self, data: str) -> bytes:
        enc_key, iv, ciphertext = data.split(SEPARATOR)
        shared_key = self.private_key.decrypt(self.str_to_bytes(enc_key), padding.PKCS1v15())
        dec = Cipher(algorithms.AES(shared_key), modes.CTR(self.str_to_bytes(iv)), backend=default_backend()).decryptor()
        return dec.update(self.str_to_bytes(ciphertext)) + dec.finalize()


def create_self_signed_certificate(certfile, keyfile, hosts=("127.0.0.1", "localhost"), days: int = 365, bits: int = 2048) -> None:
    """A self-signed server certificate (subject alternative names ``hosts``) for demo networks and tests; production
    servers get theirs from a CA.  The certificate doubles as the ``ca_file`` the clients verify against."""
    import datetime
    import ipaddress

    from cryptography import x509
    from cryptography.hazmat.primitives import hashes
    from cryptography.x509.oid import NameOID

    key = rsa.generate_private_key(public_exponent=65537, key_size=bits, backend=default_backend())
    name = x509.Name([x509.NameAttribute(NameOID.COMMON_NAME, "vantage6-b200 server")])
    now = datetime.datetime.now(datetime.timezone.utc)
    sans = []
    for h in hosts:
        try:
            sans.append(x509.IPAddress(ipaddress.ip_address(h)))
        except ValueError:
            sans.append(x509.DNSName(h))
    cert = (x509.CertificateBuilder().subject_name(name).issuer_name(name).public_key(key.public_key())
            .serial_number(x509.random_serial_number()).not_valid_before(now - datetime.timedelta(minutes=5))
            .not_valid_after(now + datetime.timedelta(days=days))
            .add_extension(x509.SubjectAlternativeName(sans), critical=False)
            .add_extension(x509.BasicConstraints(ca=True, path_length=None), critical=True)
            .sign(key, hashes.SHA256()))
    Path(certfile).parent.mkdir(parents=True, exist_ok=True)
    Path(certfile).write_bytes(cert.public_bytes(serialization.Encoding.PEM))
    Path(keyfile).write_bytes(key.private_bytes(encoding=serialization.Encoding.PEM, format=serialization.PrivateFormat.TraditionalOpenSSL,
                                                encryption_algorithm=serialization.NoEncryption()))
    try:
        os.chmod(keyfile, 0o600)
    except OSError:
        pass


__all__ = ["CryptorBase", "DummyCryptor", "RSACryptor", "STRING_ENCODING", "create_self_signed_certificate"]
