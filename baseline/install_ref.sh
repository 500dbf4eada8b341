#!/usr/bin/env bash
# Puts the reference's own, UNMODIFIED Python package (network/, misc/, utils/pytorch_utils.py)
# into the git-ignored baseline/_ref/ so that it travels to the GPU box with gpurun.  It is what
# tests/test_dropin_gpu.py runs on top of this repo's plugin (pytorch_points_b200._ext aliased to
# pytorch_points._ext) to prove the drop-in claim with the reference's code instead of a re-typed
# copy.  The CUDA extension sources are NOT taken: the plugin (or oracle/_ref) stands in for them.
#
# `pip install --target baseline/_ref /root/reference` is not possible in this image: setup.py
# builds three CUDA extensions, two of which do not compile against torch 2.11 without the
# forced-include shim of oracle/build_ref.sh, and install_requires names plyfile / openmesh /
# matplotlib, none of which is in the wheelhouse.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PP_REFERENCE_ROOT:-/root/reference}/pytorch_points"
OUT="$HERE/_ref/pytorch_points"
[ -d "$REF" ] || { echo "reference package not found at $REF"; exit 3; }
rm -rf "$OUT"
mkdir -p "$OUT/utils"
cp "$REF/__init__.py" "$OUT/"
cp -r "$REF/network" "$REF/misc" "$OUT/"
cp "$REF/utils/__init__.py" "$REF/utils/pytorch_utils.py" "$OUT/utils/"
find "$OUT" -name __pycache__ -type d -prune -exec rm -rf {} +
echo "installed reference python package into $OUT"
