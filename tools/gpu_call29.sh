#!/bin/bash
timeout 300 python -m pytest tests/test_golden_gpu.py -q -m gpu -k "scaling or remap_golden" 2>&1 | tail -3
