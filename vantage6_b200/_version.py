"""Version of vantage6_b200, built the same way the reference builds its PEP-440 string
(reference vantage6/cli/_version.py:10-20): ``version_info = (major, minor, patch, stage, build, post)``."""
import json
import os

here = os.path.abspath(os.path.dirname(__file__))
with open(os.path.join(here, "__build__")) as fp:
    __build__ = json.load(fp)

version_info = (3, 1, 0, "final", __build__, 0)
_specifier_ = {"alpha": "a", "beta": "b", "candidate": "rc", "final": ""}
version = f"{version_info[0]}.{version_info[1]}.{version_info[2]}"
pre_release = "" if version_info[3] == "final" else "." + _specifier_[version_info[3]] + str(version_info[4])
post_release = "" if not version_info[5] else f".post{version_info[5]}"
__version__ = f"{version}{pre_release}{post_release}"
