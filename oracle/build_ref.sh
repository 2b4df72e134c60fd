#!/usr/bin/env bash
# Build the UNMODIFIED reference engine (cythonsim/main.pyx + simrandom.pyx) into oracle/_ref/.
#
# TEST INFRASTRUCTURE ONLY.  Nothing under reina_b200/ may import oracle/_ref; it is used by
# tests/golden/make_golden.py (golden ensemble statistics), tests, and bench.py --impl reference /
# the cpu_baseline leg.
#
# The reference sources are read where they lie under $REF (default /root/reference); nothing is
# copied into the repo.  Generated C goes to a temp dir, only the two extension modules land in
# oracle/_ref/cythonsim/.  Recipe follows SURVEY.md Appendix A:
#   cython -3 -X legacy_implicit_noexcept=True   (restores the pinned cython==3.0a6 semantics)
#   gcc -O2 -fopenmp (main.pyxbld:14-15), libnpyrandom for simrandom (simrandom.pyxbld:15-16)
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/cythonsim" ]; then
  echo "build_ref: $REF/cythonsim not present (GPU box?) - keeping prebuilt oracle/_ref" >&2
  exit 0
fi
EXT=$(python3 -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
if [ -f "$OUT/cythonsim/main$EXT" ] && [ -f "$OUT/cythonsim/simrandom$EXT" ] && [ "${FORCE:-0}" != 1 ]; then
  echo "build_ref: up to date"; exit 0
fi
TMP=$(mktemp -d /tmp/reina_ref_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT
mkdir -p "$OUT/cythonsim"
PYINC=$(python3 -c "import sysconfig;print(sysconfig.get_paths()['include'])")
NPINC=$(python3 -c "import numpy;print(numpy.get_include())")
# cython resolves `from cythonsim.simrandom cimport RandomPool` through the package dir in $REF
( cd "$REF" && cython -3 -X legacy_implicit_noexcept=True cythonsim/simrandom.pyx -o "$TMP/simrandom.c" \
            && cython -3 -X legacy_implicit_noexcept=True cythonsim/main.pyx -o "$TMP/main.c" ) 2> "$TMP/cython.log" \
  || { cat "$TMP/cython.log" >&2; exit 1; }
gcc -O2 -fPIC -shared -w -DNPY_NO_DEPRECATED_API -I"$PYINC" -I"$NPINC" "$TMP/simrandom.c" \
    -o "$OUT/cythonsim/simrandom$EXT" -L"$NPINC/../../random/lib" -lnpyrandom -lm
gcc -O2 -fPIC -shared -w -fopenmp -DNPY_NO_DEPRECATED_API -I"$PYINC" -I"$NPINC" "$TMP/main.c" \
    -o "$OUT/cythonsim/main$EXT" -lm
: > "$OUT/cythonsim/__init__.py"   # NOT the reference's (it triggers pyximport)
echo "build_ref: built $OUT/cythonsim/{simrandom,main}$EXT"
