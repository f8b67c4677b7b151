#!/bin/bash
# Installs the UNMODIFIED reference (uw-comphys/openccm) into baseline/_ref and a copy of the inputs of its OpenFOAM example into
# baseline/_ref_examples, so that the GPU box (which has no /root/reference) can run `openccm.run` through the drop-in dispatcher
# (tests/test_gpu_dropin.py).  Both directories are git-ignored (they are the reference's files, not product source) and are NOT
# gpurun-ignored, so they travel with the snapshot.  Runs only where /root/reference exists (the build container).
#   --no-deps: the wheelhouse of this image carries no numpy/scipy wheels (they are installed in the interpreter already);
#   the source tree is read-only, so pip builds from a copy under /tmp.
set -euo pipefail
REPO="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
[ -d "$REF/openccm" ] || { echo "no reference at $REF"; exit 0; }
TMP="$(mktemp -d)"
cp -r "$REF" "$TMP/ref"
rm -rf "$REPO/baseline/_ref"
mkdir -p "$REPO/baseline"
python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target "$REPO/baseline/_ref" "$TMP/ref"
# inputs of examples/OpenFOAM/pipe_with_recirc: mesh, fields and CONFIG (the gmsh sources of the mesh are not needed to run)
EX="$REPO/baseline/_ref_examples/pipe_with_recirc"
rm -rf "$EX"; mkdir -p "$EX"
for f in 0 constant system CONFIG; do cp -r "$REF/examples/OpenFOAM/pipe_with_recirc/$f" "$EX/"; done
cp "$REF/examples/OpenCMP/pipe_with_recirc_2d/reactions" "$EX/reactions_nacl"
rm -rf "$TMP"
echo "reference installed: $(ls "$REPO/baseline/_ref")"
