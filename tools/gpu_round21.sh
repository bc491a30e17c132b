#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "operand_column" --timeout 300 -p no:cacheprovider 2>&1 | grep -E "Error|error|assert|^E" | head -20
