#!/bin/bash
# apply_reference_patches.sh -- build the GPU-enabled Fortran hosts where a toolchain exists.
#
# Copies the handful of reference sources cdfmoc / cdfmocsig / cdfsigtrp need (src/Makefile:305-315,570-586) from a CDFTOOLS checkout
# into a build directory, applies cdftools_b200/fortran/patches/*.patch (the call-site changes of INTEGRATION.md), adds
# cdftools_b200/fortran/{cdfgpu_mod,cdfio_pinned}.f90 and links against libcdfgpu.so.  The result is the reference's own
# cdfmoc / cdfmocsig -- same command line, same cdfio, same output block -- with the hot loop nests on the GPU.
#
#   tools/apply_reference_patches.sh /path/to/CDFTOOLS [builddir]
#   env: FC (gfortran), NETCDF_INC / NETCDF_LIBS (default: nf-config), FFLAGS_EXTRA (e.g. "-D key_netcdf4" for a real -nc4)
set -euo pipefail
REF=${1:?usage: apply_reference_patches.sh /path/to/CDFTOOLS [builddir]}
HERE=$(cd "$(dirname "$0")/.." && pwd)
BUILD=${2:-$HERE/build/fortran_host}
FC=${FC:-gfortran}
command -v "$FC" >/dev/null || { echo "no Fortran compiler ($FC)"; exit 3; }
if command -v nf-config >/dev/null; then
  NETCDF_INC=${NETCDF_INC:-$(nf-config --includedir)}
  NETCDF_LIBS=${NETCDF_LIBS:-$(nf-config --flibs)}
else
  : "${NETCDF_INC:?set NETCDF_INC}" "${NETCDF_LIBS:?set NETCDF_LIBS}"
fi
[ -f "$HERE/cdftools_b200/libcdfgpu.so" ] || python3 -m cdftools_b200.build
mkdir -p "$BUILD/src"
cp "$REF"/src/{modcdfnames.F90,modcdfnames_CMIP6.h90,cdfio.F90,eos.f90,cdftools.f90,modutils.f90,cdfmoc.f90,cdfmocsig.f90,cdfsigtrp.f90} "$BUILD/src/"
( cd "$BUILD" && for p in "$HERE"/cdftools_b200/fortran/patches/*.patch; do patch -p1 < "$p"; done )
cp "$HERE"/cdftools_b200/fortran/{cdfgpu_mod,cdfio_pinned}.f90 "$BUILD/src/"
FFLAGS="-O -I$NETCDF_INC -fno-second-underscore -ffree-line-length-256 ${FFLAGS_EXTRA:-}"
GPU="-L$HERE/cdftools_b200 -lcdfgpu -Wl,-rpath,$HERE/cdftools_b200"
( cd "$BUILD/src"
  for f in modcdfnames.F90 cdfio.F90 eos.f90 cdftools.f90 modutils.f90 cdfgpu_mod.f90 cdfio_pinned.f90; do $FC -c $f $FFLAGS; done
  $FC cdfmoc.f90    -o ../cdfmoc    cdfio.o eos.o modcdfnames.o cdftools.o cdfgpu_mod.o cdfio_pinned.o $FFLAGS $NETCDF_LIBS $GPU
  $FC cdfmocsig.f90 -o ../cdfmocsig cdfio.o eos.o modcdfnames.o modutils.o cdfgpu_mod.o cdfio_pinned.o $FFLAGS $NETCDF_LIBS $GPU
  $FC cdfsigtrp.f90 -o ../cdfsigtrp cdfio.o eos.o modcdfnames.o modutils.o cdfgpu_mod.o $FFLAGS $NETCDF_LIBS $GPU )
echo "built $BUILD/cdfmoc, $BUILD/cdfmocsig and $BUILD/cdfsigtrp (GPU-enabled Fortran hosts)"
