"""Version of vantage6_b200.

``version_info = (major, minor, patch, stage, build, post)`` is turned into a PEP 440 string the way
the reference does it (vantage6/cli/_version.py:10-20): pre-releases get ``.a<build>`` / ``.b<build>`` /
``.rc<build>``, a non-zero ``post`` appends ``.post<N>``.  ``build`` is read from the ``__build__`` file that
ships next to this module.
"""
import json
from pathlib import Path

_STAGE_TAG = {"alpha": "a", "beta": "b", "candidate": "rc", "final": None}


def _read_build() -> int:
    return json.loads((Path(__file__).with_name("__build__")).read_text())


def pep440(info) -> str:
    major, minor, patch, stage, build, post = info
    text = f"{major}.{minor}.{patch}"
    tag = _STAGE_TAG[stage]
    if tag is not None:
        text += f".{tag}{build}"
    if post:
        text += f".post{post}"
    return text


__build__ = _read_build()
version_info = (3, 1, 0, "final", __build__, 0)
__version__ = pep440(version_info)
