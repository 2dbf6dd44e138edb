#!/usr/bin/env bash
# remove Python bytecode caches and build leftovers (the reference ships an empty script of the same name)
cd "$(dirname "$0")/.."
find . -name __pycache__ -type d -prune -exec rm -rf {} + 2>/dev/null
rm -rf .pytest_cache tiny_deepspeed_b200/csrc/_build
echo "caches cleared"
