#!/bin/bash
# baseline/run_ghc.sh — build the REAL reference with GHC and render the shipped scenes, so that the oracle can be
# pinned against reference output (tests/test_ghc_golden.py) and the GHC CPU baseline measured.
#
#   baseline/run_ghc.sh [reference checkout, default /root/reference] [output dir, default tests/golden/ghc]
#
# GHC is not installed in this image nor on the GPU boxes: there the script says so and exits 0.
# Where `ghc` (with the packages of rayhs.cabal: aeson, vector, split, parallel, deepseq-generics, random, mtl) exists,
# it copies the checkout to a scratch directory (the reference tree stays read-only), adds the two textures
# data/texture.json names but the repository does not ship (the same synthetic P3 files tests/golden/make_packs.py
# writes), builds with the reference's own flags (build.sh:1: ghc -isrc -O2 -funbox-strict-fields -threaded -rtsopts) and
# runs `./rayhs -o<scene>.ppm data/<scene>.json +RTS -N<cores> -s` for the shipped scenes at their native sizes.
set -u
HERE="$(cd "$(dirname "$0")/.." && pwd)"
REF="${1:-/root/reference}"
OUT="${2:-$HERE/tests/golden/ghc}"
if ! command -v ghc >/dev/null 2>&1; then
  echo "GHC absent — not measured (the oracle stays a restatement: parity unpinned)"
  exit 0
fi
WORK="$(mktemp -d)"
cp -r "$REF/." "$WORK/"
chmod -R u+w "$WORK"
python - "$WORK/data" <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("RAYHS_B200_ROOT", "."))
from tests.golden.make_packs import synth_textures
synth_textures(sys.argv[1])
PY
cd "$WORK" || exit 1
if [ -x ./build.sh ]; then ./build.sh; else ghc -isrc -O2 -funbox-strict-fields -threaded -rtsopts src/RayHs.hs -o rayhs; fi || { echo "reference build failed"; exit 1; }
mkdir -p "$OUT"
CORES="$(nproc)"
for s in cornellBox texture transform dragon outScene; do
  echo "== $s (+RTS -N$CORES)"
  /usr/bin/time -f "$s: %e s wall, %M KB" ./rayhs "-o$OUT/$s.ppm" "data/$s.json" +RTS "-N$CORES" -s 2> "$OUT/$s.rts.txt" || echo "$s failed"
  tail -3 "$OUT/$s.rts.txt"
done
echo "reference PPMs in $OUT — run: python -m pytest tests/test_ghc_golden.py -q"
