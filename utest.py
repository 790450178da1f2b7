#!/usr/bin/env python3
"""unittest-style runner (parity: reference utest.py:1-12); the primary runner is pytest."""
from pathlib import Path

from vantage6_b200.common.utest import find_tests, run_tests


def run():
    run_tests(find_tests(str(Path(__file__).parent / "tests")))


if __name__ == "__main__":
    run()
