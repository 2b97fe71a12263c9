#!/bin/bash
export PYTHONUNBUFFERED=1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
timeout 300 python scripts/debug_image_grads.py 2>&1 | tail -12
