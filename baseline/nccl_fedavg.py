#!/usr/bin/env python
"""Convenience launcher for the NCCL baseline arm: identical to `bench.py --impl nccl`."""
import os
import runpy
import sys

if __name__ == "__main__":
    sys.argv = [os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), "--impl", "nccl"] + sys.argv[1:]
    runpy.run_path(sys.argv[0], run_name="__main__")
