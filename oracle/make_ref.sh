#!/usr/bin/env bash
# Stages the UNMODIFIED reference (tdunnlab/scrubvae, pure Python, `dependencies = []`) into oracle/_ref/ so that it
# travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored).  Nothing is copied into
# tracked files: the package is installed by pip from a scratch copy of the tree (the build writes egg-info into the
# source directory and /root/reference is read-only).  Used by
#   * bench.py --impl reference            (the reference's own train_test_epoch on the host cores)
#   * bench.py gpu_eager_baseline          (the same function on cuda:0 — PyTorch-eager, the same-box GPU comparator)
#   * tests/test_reference_gpu.py          (B=2048 parity against the reference run on the same GPU)
# Re-run whenever /root/reference changes:   bash oracle/make_ref.sh
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${SCV_REFERENCE_ROOT:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src/scrubvae" ]; then
  echo "make_ref: no reference tree at $REF (GPU box?): keeping whatever is staged in $OUT" >&2
  exit 0
fi
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$TMP/ref"
cp -r "$REF/pyproject.toml" "$REF/src" "$TMP/ref/"
[ -f "$REF/README.md" ] && cp "$REF/README.md" "$TMP/ref/"
[ -f "$REF/LICENSE" ] && cp "$REF/LICENSE" "$TMP/ref/"
rm -rf "$OUT"
mkdir -p "$OUT"
PY="${PYTHON:-python}"
if ! "$PY" -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
      --target "$OUT" "$TMP/ref" >"$TMP/pip.log" 2>&1; then
  echo "make_ref: pip install failed, staging the package directory as it lies" >&2
  cat "$TMP/pip.log" >&2
  cp -r "$REF/src/scrubvae" "$OUT/scrubvae"
fi
# the skeleton definition the reference reads through neuroposelib.read.config (kinematic tree, offsets)
mkdir -p "$OUT/configs"
cp "$REF/configs/mouse_skeleton.yaml" "$OUT/configs/"
find "$OUT" -name '__pycache__' -type d -prune -exec rm -rf {} +
( cd "$REF" && find src/scrubvae -name '*.py' -print0 | sort -z | xargs -0 sha256sum ) > "$OUT/SOURCE_SHA256"
echo "make_ref: staged $(find "$OUT/scrubvae" -name '*.py' | wc -l) reference modules into $OUT"
