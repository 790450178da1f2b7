"""Tiny validation library with the slice of the ``schema`` package API that the reference's
configuration classes use (reference vantage6/cli/configuration_manager.py:1,11-54):
``And``, ``Or``, ``Use``, ``Optional`` and dict / callable / type / literal specs.
(``schema`` itself is not installed in this environment.)"""
from __future__ import annotations

from typing import Any


class SchemaError(Exception):
    pass


class Use:
    def __init__(self, fn):
        self.fn = fn

    def validate(self, data):
        try:
            return self.fn(data)
        except Exception as e:  # noqa: BLE001
            raise SchemaError(f"{self.fn.__name__ if hasattr(self.fn, '__name__') else self.fn}({data!r}) raised {e!r}")


class And:
    def __init__(self, *specs):
        self.specs = specs

    def validate(self, data):
        for s in self.specs:
            data = _validate(s, data)
        return data


class Or:
    def __init__(self, *specs):
        self.specs = specs

    def validate(self, data):
        errors = []
        for s in self.specs:
            try:
                return _validate(s, data)
            except SchemaError as e:
                errors.append(str(e))
        raise SchemaError(f"{data!r} did not match any alternative: {errors}")


class Optional:
    def __init__(self, key):
        self.key = key

    def __hash__(self):
        return hash(("optional", self.key))

    def __eq__(self, other):
        return isinstance(other, Optional) and other.key == self.key


def _validate(spec: Any, data: Any) -> Any:
    if isinstance(spec, (Use, And, Or)):
        return spec.validate(data)
    if isinstance(spec, dict):
        if not isinstance(data, dict):
            raise SchemaError(f"{data!r} should be a mapping")
        out = dict(data)
        literal = {k: v for k, v in spec.items() if isinstance(k, (str, Optional))}
        generic = [(k, v) for k, v in spec.items() if not isinstance(k, (str, Optional))]
        seen = set()
        for k, sub in literal.items():
            name = k.key if isinstance(k, Optional) else k
            if name not in data:
                if isinstance(k, Optional):
                    continue
                raise SchemaError(f"Missing key: {name!r}")
            out[name] = _validate(sub, data[name])
            seen.add(name)
        for dk, dv in data.items():
            if dk in seen:
                continue
            for gk, gv in generic:
                try:
                    nk = _validate(gk, dk)
                    out.pop(dk, None)
                    out[nk] = _validate(gv, dv)
                    break
                except SchemaError:
                    continue
            # unknown extra keys are tolerated (vantage6 configs carry optional keys such as
            # `image`, `vpn_subnet`, `jwt_secret_key`, `rabbitmq_uri`)
        return out
    if isinstance(spec, type):
        if not isinstance(data, spec):
            raise SchemaError(f"{data!r} should be instance of {spec.__name__!r}")
        return data
    if callable(spec):
        try:
            ok = spec(data)
        except Exception as e:  # noqa: BLE001
            raise SchemaError(f"{spec}({data!r}) raised {e!r}")
        if not ok:
            raise SchemaError(f"{getattr(spec, '__name__', spec)}({data!r}) should evaluate to True")
        return data
    if spec is None:
        if data is not None:
            raise SchemaError(f"{data!r} should be None")
        return data
    if spec != data:
        raise SchemaError(f"{data!r} does not match {spec!r}")
    return data


class Schema:
    def __init__(self, spec, ignore_extra_keys: bool = True):
        self.spec = spec

    def validate(self, data):
        return _validate(self.spec, data)

    def is_valid(self, data) -> bool:
        try:
            self.validate(data)
            return True
        except SchemaError:
            return False
