#!/usr/bin/env python3
"""Run the test-suite with the stdlib ``unittest`` machinery (the reference ships the same entry point,
utest.py:1-12).  ``python -m pytest tests`` is the primary runner; this one needs no third-party runner."""
import sys
from pathlib import Path

from vantage6_b200.common.utest import find_tests, run_tests

TESTS = Path(__file__).resolve().parent / "tests"


def run() -> bool:
    return run_tests(find_tests(str(TESTS)))


if __name__ == "__main__":
    sys.exit(0 if run() in (None, True) else 1)
