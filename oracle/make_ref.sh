#!/bin/bash
# TEST INFRASTRUCTURE.  Stages the UNMODIFIED reference package next to the
# oracle so that the numba + multiprocessing CPU arm of bench.py
# (`--impl reference`, kind "reference") can run on the GPU box, where
# /root/reference does not exist.  Output only into oracle/_ref/ (ignored by
# git, shipped to the box by gpurun like the built .so files).  Nothing under
# ngmix_b200/ ever imports it.
#
#   oracle/_ref/ngmix/            copy of /root/reference/ngmix (python only)
#   oracle/_ref/galsim/__init__.py  EMPTY stub: ngmix/__init__.py imports
#       ngmix.metacal, which imports galsim at module level; no galsim
#       function is called by the fitting path that is timed.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/ngmix" ]; then
    echo "make_ref.sh: $SRC/ngmix not found (nothing staged)"; exit 0
fi
rm -rf "$HERE/_ref"
mkdir -p "$HERE/_ref/galsim"
cp -r "$SRC/ngmix" "$HERE/_ref/ngmix"
find "$HERE/_ref" -name "__pycache__" -type d -prune -exec rm -rf {} +
rm -rf "$HERE/_ref/ngmix/tests"
: > "$HERE/_ref/galsim/__init__.py"
echo "staged reference ngmix $(grep -o "'[0-9][^']*'" "$HERE/_ref/ngmix/_version.py" 2>/dev/null | head -1) in $HERE/_ref"
