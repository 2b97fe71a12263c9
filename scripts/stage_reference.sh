#!/bin/bash
# Stage the unmodified reference under baseline/_ref (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the
# built .so files, never into history).  The reference has no setup.py, so `pip install --target baseline/_ref` does not
# apply; a plain copy is the equivalent.  Used by tests/test_gpu_reference_drivers.py (unmodified
# text.py / image.py against both back-ends on the same GPU) — skipped when absent.
set -e
cd "$(dirname "$0")/.."
src="${VAE_REF_PATH:-/root/reference}"
[ -d "$src/modules" ] || { echo "no reference at $src"; exit 1; }
rm -rf baseline/_ref
mkdir -p baseline/_ref
cp -r "$src"/{text.py,image.py,toy.py,logger.py,modules,data,config} baseline/_ref/
find baseline/_ref -name __pycache__ -prune -exec rm -rf {} +
echo "staged $(find baseline/_ref -name '*.py' | wc -l) files under baseline/_ref"
