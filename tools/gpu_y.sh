#!/bin/bash
# the goldens added after the last full run, on the GPU
timeout 600 python -m pytest tests/test_golden.py -m gpu -q 2>&1 | tail -3 | cut -c1-200
