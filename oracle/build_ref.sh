#!/bin/bash
# Recipe for oracle/_ref: the UNMODIFIED reference package (lucaslingle/pytorch_ddp_resnet), installed from
# /root/reference into oracle/_ref/ so that bench.py's reference arm, its cpu_baseline and its gpu_reference
# block can run the reference's own code (resnet.algos.training.training_loop, resnet.architectures.resnet.ResNet)
# on the GPU box, where /root/reference does not exist. oracle/_ref/ is a build artefact: git-ignored (never
# committed), not gpurun-ignored (it travels with the snapshot). Test infrastructure only.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
[ -d "$SRC/resnet" ] || { echo "build_ref: $SRC/resnet not found (nothing to do on this box)"; exit 0; }
rm -rf "$DST" && mkdir -p "$DST"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/src"          # the source tree is read-only: build from a copy
python -m pip install --quiet --no-index --no-deps --no-build-isolation --target "$DST" "$TMP/src" \
    > "$TMP/pip.log" 2>&1 || echo "build_ref: pip install failed (see below), falling back to a plain package copy"
if [ ! -f "$DST/resnet/architectures/resnet.py" ]; then
    # setup.py declares the package through py_modules=["resnet"], which installs no sub-packages
    tail -3 "$TMP/pip.log" || true
    rm -rf "$DST/resnet"
    cp -r "$SRC/resnet" "$DST/resnet"
fi
find "$DST" -name __pycache__ -type d -prune -exec rm -rf {} +
rm -rf "$TMP"
python - <<PY
import sys
sys.path.insert(0, "$DST")
import resnet.architectures.resnet, resnet.algos.training  # noqa
print("build_ref: oracle/_ref ready:", resnet.__file__ if hasattr(resnet, "__file__") and resnet.__file__ else "$DST/resnet")
PY
