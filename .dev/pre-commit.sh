#!/usr/bin/env bash
# Formatting / lint gate (counterpart of the reference's .dev/pre-commit.sh): run from the repo root.
# Tools are optional in the offline image; each step is skipped with a note when its tool is missing.
set -u
cd "$(dirname "$0")/.."
status=0
run() { if command -v "$1" >/dev/null 2>&1; then echo "== $*"; "$@" || status=1; else echo "== skip: $1 not installed"; fi; }
run isort --check-only --diff distribuuuu_b200 tests tools tutorial train_net.py test_net.py bench.py
run black --check --line-length 120 distribuuuu_b200 tests tools tutorial train_net.py test_net.py bench.py
run flake8 distribuuuu_b200 tests tools tutorial train_net.py test_net.py bench.py
# always available: byte-compile everything and make sure the CPU test-suite collects
python -m compileall -q distribuuuu_b200 tests tools tutorial train_net.py test_net.py bench.py __graft_entry__.py || status=1
python -m pytest tests -q -m "not gpu" --collect-only >/dev/null || status=1
exit $status
