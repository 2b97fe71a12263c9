#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_image.py -m gpu -q -k "generation or ancestral" --timeout 600 2>&1 | grep -v "^E   *+" | tail -40 | cut -c1-400
