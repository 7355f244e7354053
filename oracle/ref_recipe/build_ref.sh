#!/bin/sh
# Builds the UNMODIFIED reference (perazz/fitpack) hot path from the sources where they lie into
# oracle/_ref/ (git-ignored; travels to the GPU box like the other built .so files):
#   libfitpack_ref.so     -O2 -ffp-contract=off  -> golden vectors (oracle/ref_recipe/make_ref_golden.py)
#   libfitpack_ref_O3.so  -O3 -ffp-contract=off  -> CPU baseline of bench.py (kind "reference")
# Only two source files are needed: the numerical core and its bind(C) layer, which exports the same flat
# symbols as include/fitpack_b200.h (fp_curfit_c, fp_splev_c, fp_percur_c, fp_parcur_c, fp_curev_c,
# fp_regrid_c, fp_bispev_c, fp_splder_c, ...).  No fpm, no cmake.  This image has NO Fortran compiler, so
# here the script reports that and exits 3; on a box with gfortran it closes the "parity unpinned" gap
# (SURVEY 8c step 5): run it, then `python oracle/ref_recipe/make_ref_golden.py`, commit tests/golden/ref_*.npz.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${FITPACK_REF:-/root/reference}
FC=${FC:-gfortran}
if ! command -v "$FC" >/dev/null 2>&1; then
    echo "build_ref.sh: no Fortran compiler ($FC) in this image: oracle/_ref not built" >&2
    exit 3
fi
if [ ! -f "$REF/src/fitpack_core.F90" ]; then
    echo "build_ref.sh: reference sources not found under $REF" >&2
    exit 4
fi
OUT="$ROOT/oracle/_ref"
mkdir -p "$OUT/mod_O2" "$OUT/mod_O3"
COMMON="-cpp -fPIC -shared -ffp-contract=off -fno-fast-math -fopenmp"
"$FC" -O2 $COMMON -J "$OUT/mod_O2" -o "$OUT/libfitpack_ref.so" \
    "$REF/src/fitpack_core.F90" "$REF/src/fitpack_core_c.f90"
"$FC" -O3 $COMMON -J "$OUT/mod_O3" -o "$OUT/libfitpack_ref_O3.so" \
    "$REF/src/fitpack_core.F90" "$REF/src/fitpack_core_c.f90"
echo "$OUT/libfitpack_ref.so"
