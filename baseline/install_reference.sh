#!/usr/bin/env bash
# Installs the UNMODIFIED reference (BIGBALLON/distribuuuu) into baseline/_ref (git-ignored).
# /root/reference has neither setup.py nor pyproject.toml, so `pip install /root/reference` fails with
# "Neither 'setup.py' nor 'pyproject.toml' found"; we therefore install from a /tmp copy that only ADDS a
# three-line setup.py (no source file is touched; `diff -r` against /root/reference/distribuuuu is clean).
set -euo pipefail
cd "$(dirname "$0")/.."
rm -rf /tmp/refpkg baseline/_ref
mkdir -p /tmp/refpkg
cp -r /root/reference/. /tmp/refpkg/
cat > /tmp/refpkg/setup.py <<'PY'
from setuptools import find_packages, setup
setup(name="distribuuuu", version="1.0.0", packages=find_packages(include=["distribuuuu", "distribuuuu.*"]))
PY
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target baseline/_ref /tmp/refpkg
diff -r -x __pycache__ /root/reference/distribuuuu baseline/_ref/distribuuuu && echo "reference installed unmodified"
