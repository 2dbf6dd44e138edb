#!/usr/bin/env bash
# Installs the UNMODIFIED reference into baseline/_ref (git-ignored) for `bench.py --impl reference`.
# /root/reference has no setup.py/pyproject (pip refuses it), so a copy under /tmp gets a packaging-only
# setup.py; no source file of the reference is changed (verified with diff -r below).
set -euo pipefail
cd "$(dirname "$0")/.."
rm -rf baseline/_ref /tmp/ref_src
cp -r /root/reference /tmp/ref_src
cat > /tmp/ref_src/setup.py <<'PY'
from setuptools import setup, find_namespace_packages
setup(name="tiny_deepspeed_reference", version="0.0.0",
      packages=find_namespace_packages(include=["tiny_deepspeed*", "example*"]))
PY
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/ref_src
find baseline/_ref -name __pycache__ -prune -exec rm -rf {} \;
diff -rq /root/reference/tiny_deepspeed baseline/_ref/tiny_deepspeed && diff -rq /root/reference/example baseline/_ref/example && echo "reference installed, sources identical"
