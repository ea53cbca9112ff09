#!/usr/bin/env bash
# Offline install of the UNMODIFIED reference into baseline/_ref (git-ignored; travels with gpurun snapshots).
#
# The reference's setup.py extracts its native kernels from a vLLM wheel (setup.py:186-215); with no network we
# hand it, through its own GLLM_PRECOMPILED_WHEEL_LOCATION switch, a local wheel-shaped zip whose members are
# placeholders, and then point those placeholders at the vLLM libraries already installed in this image
# (symlinks: ~1 GB of .so files do not have to travel). Dependencies are not resolved (--no-deps): torch /
# transformers of the image are used as they are, the PyPI `logger` package is provided by baseline/shims/.
# NOTE: the image has vLLM 0.22, the reference pins 0.11 — whether the op signatures still match is checked on the
# GPU by baseline/run_reference.py; bench.py falls back to {"unavailable": ...} when they do not.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref=${REFERENCE_SRC:-/root/reference}
vllm_dir=$(python -c "import vllm, os; print(os.path.dirname(vllm.__file__))" 2>/dev/null | tail -1)
work=$(mktemp -d /tmp/gllm_ref_build.XXXXXX)
trap 'rm -rf "$work"' EXIT
cp -r "$ref" "$work/src"
python - "$work" "$vllm_dir" <<'PY'
import os, sys, zipfile
work, vllm_dir = sys.argv[1], sys.argv[2]
members = ["vllm/_C.abi3.so", "vllm/_moe_C.abi3.so", "vllm/_flashmla_C.abi3.so", "vllm/_flashmla_extension_C.abi3.so",
           "vllm/vllm_flash_attn/_vllm_fa2_C.abi3.so", "vllm/vllm_flash_attn/_vllm_fa3_C.abi3.so",
           "vllm/cumem_allocator.abi3.so"]
with zipfile.ZipFile(os.path.join(work, "vllm-local-cp38-abi3-linux_x86_64.whl"), "w") as z:
    for m in members:
        if os.path.exists(os.path.join(os.path.dirname(vllm_dir), m)):
            z.writestr(m, b"placeholder: replaced by a symlink to the installed vLLM library\n")
    fa = os.path.join(vllm_dir, "vllm_flash_attn")
    for root, _, files in os.walk(fa):
        for f in files:
            if f.endswith(".py"):
                full = os.path.join(root, f)
                z.write(full, os.path.relpath(full, os.path.dirname(vllm_dir)))
PY
rm -rf "$here/_ref"
( cd "$work/src" && GLLM_PRECOMPILED_WHEEL_LOCATION="$work/vllm-local-cp38-abi3-linux_x86_64.whl" \
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$here/_ref" . )
# the reference is meant for `pip install -e .`: its find_packages() skips the sub-packages without __init__.py
# (layers/, models/, entrypoints/, ...) — complete the tree with the untouched sources
cp -rn "$ref/gllm/." "$here/_ref/gllm/"
# setup.py registers the extracted files as package data of "vllm" (not "gllm"), so a non-editable install does
# not carry them: place them next to the installed sources ourselves — the flash-attn python wrappers it
# extracted into the build tree, and symlinks to the image's vLLM libraries for the native modules
# (same absolute path on every box of this image).
if [[ -d "$work/src/gllm/vllm_flash_attn" ]]; then
  mkdir -p "$here/_ref/gllm/vllm_flash_attn"
  ( cd "$work/src/gllm/vllm_flash_attn" && find . -name "*.py" -exec cp --parents {} "$here/_ref/gllm/vllm_flash_attn/" \; )
fi
for rel in _C.abi3.so _moe_C.abi3.so _flashmla_C.abi3.so _flashmla_extension_C.abi3.so cumem_allocator.abi3.so \
           vllm_flash_attn/_vllm_fa2_C.abi3.so vllm_flash_attn/_vllm_fa3_C.abi3.so; do
  if [[ -f "$vllm_dir/$rel" ]]; then
    mkdir -p "$(dirname "$here/_ref/gllm/$rel")"
    ln -sf "$vllm_dir/$rel" "$here/_ref/gllm/$rel"
  fi
done
ls -la "$here/_ref/gllm/"*.so "$here/_ref/gllm/vllm_flash_attn/" | head -30
echo "reference installed into $here/_ref"
