#!/bin/bash
# make_reference_golden.sh -- pin the oracle (and the CUDA path) on the REAL Fortran reference.
#
# The build image has no Fortran compiler and no NetCDF library, so parity is "unpinned" there.
# On any machine with gfortran + netcdf-fortran this script
#   1. compiles cdfmoc, cdfmocsig and cdfsigtrp from the unmodified reference sources (the few files src/Makefile:296-315,570-586
#      list for them; the reference's own build system is not run),
#   2. writes the synthetic input files of tests/ (cdftools_b200/ncfiles.py: seeded, identical on every machine),
#   3. runs the reference tools on them (-decomp, -isodep, -r 2000, -full, ... : the cases of tools/reference_cases.json),
#   4. stores the outputs under tests/golden/ref/<case>/ together with the command line that made them.
# tests/test_reference_golden.py then compares the oracle AND (on a GPU box) the CUDA command-line twins with those
# files, variable by variable; it is skipped while tests/golden/ref/ holds no case.
#
#   tools/make_reference_golden.sh /path/to/CDFTOOLS [outdir]
#   env: FC (default gfortran), NETCDF_INC / NETCDF_LIB (default: nf-config), FFLAGS_EXTRA
set -euo pipefail
REF=${1:?usage: make_reference_golden.sh /path/to/CDFTOOLS [outdir]}
HERE=$(cd "$(dirname "$0")/.." && pwd)
OUT=${2:-$HERE/tests/golden/ref}
FC=${FC:-gfortran}
command -v "$FC" >/dev/null || { echo "no Fortran compiler ($FC): the reference cannot be built here"; exit 3; }
if command -v nf-config >/dev/null; then
  NETCDF_INC=${NETCDF_INC:-$(nf-config --includedir)}
  NETCDF_LIBS=${NETCDF_LIBS:-$(nf-config --flibs)}
else
  NETCDF_INC=${NETCDF_INC:?set NETCDF_INC (directory of netcdf.mod)}
  NETCDF_LIBS=${NETCDF_LIBS:--L${NETCDF_LIB:?set NETCDF_LIB} -lnetcdff -lnetcdf}
fi
BUILD=$(mktemp -d)
trap 'rm -rf "$BUILD"' EXIT
# the reference's own flags (Macrolib/macro.gfortran:25): -O, no key_netcdf4 needed for classic files
FFLAGS="-O -I$NETCDF_INC -fno-second-underscore -ffree-line-length-256 ${FFLAGS_EXTRA:-}"
cp "$REF"/src/{modcdfnames.F90,modcdfnames_CMIP6.h90,cdfio.F90,eos.f90,cdftools.f90,modutils.f90,cdfmoc.f90,cdfmocsig.f90,cdfsigtrp.f90} "$BUILD"/
( cd "$BUILD"
  for f in modcdfnames.F90 cdfio.F90 eos.f90 cdftools.f90 modutils.f90; do $FC -c $f $FFLAGS; done
  $FC cdfmoc.f90 -o cdfmoc cdfio.o eos.o modcdfnames.o cdftools.o $FFLAGS $NETCDF_LIBS
  $FC cdfmocsig.f90 -o cdfmocsig cdfio.o eos.o modcdfnames.o modutils.o $FFLAGS $NETCDF_LIBS
  $FC cdfsigtrp.f90 -o cdfsigtrp cdfio.o eos.o modcdfnames.o modutils.o $FFLAGS $NETCDF_LIBS )
echo "built $BUILD/cdfmoc, $BUILD/cdfmocsig and $BUILD/cdfsigtrp with $($FC --version | head -1)"
python3 "$HERE/tools/reference_cases.py" run --cdfmoc "$BUILD/cdfmoc" --cdfmocsig "$BUILD/cdfmocsig" --cdfsigtrp "$BUILD/cdfsigtrp" --out "$OUT" \
        --note "$($FC --version | head -1); FFLAGS=$FFLAGS"
echo "reference outputs written under $OUT"
