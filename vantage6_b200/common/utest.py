"""Test discovery / running helpers -- the ``vantage6.common.utest`` contract (``find_tests(path)``,
``run_tests(suites)``) the reference's ``utest.py`` is written against (reference utest.py:3-8).

The suites are stdlib ``unittest`` suites.  This repository's own tests are pytest-style functions, which the
stdlib loader does not see: :func:`find_tests` remembers where it looked and :func:`run_tests` hands an empty
discovery over to pytest (CPU tests only), so ``python utest.py`` runs the real test-suite either way.
"""
from __future__ import annotations

import sys
import unittest
from typing import Optional


class DiscoveredSuite(unittest.TestSuite):
    """A test suite that knows the directory it was discovered in."""

    origin: Optional[str] = None


def find_tests(path: str = None, pattern: str = "test_*.py") -> DiscoveredSuite:
    where = path or "."
    suite = DiscoveredSuite()
    suite.addTests(unittest.TestLoader().discover(where, pattern=pattern))
    suite.origin = where
    return suite


def run_tests(suites, verbosity: int = 1) -> bool:
    if suites.countTestCases() == 0 and getattr(suites, "origin", None):
        import pytest                                      # pytest-style tests: delegate

        ok = pytest.main([suites.origin, "-q", "-m", "not gpu"]) == 0
    else:
        ok = unittest.TextTestRunner(verbosity=verbosity).run(suites).wasSuccessful()
    if not ok:
        sys.exit(1)
    return ok
