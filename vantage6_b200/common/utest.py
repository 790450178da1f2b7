"""unittest discovery helpers (the ``vantage6.common.utest`` contract used by the reference's
test runner: reference utest.py:3-8 -- ``find_tests(path)`` and ``run_tests(suites)``)."""
from __future__ import annotations

import sys
import unittest


def find_tests(path: str = None, pattern: str = "test_*.py"):
    loader = unittest.TestLoader()
    return loader.discover(path or ".", pattern=pattern)


def run_tests(suites, verbosity: int = 1) -> bool:
    runner = unittest.TextTestRunner(verbosity=verbosity)
    result = runner.run(suites)
    ok = result.wasSuccessful()
    if not ok:
        sys.exit(1)
    return ok
