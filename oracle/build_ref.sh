#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Compiles the UNMODIFIED reference hot-path sources
# where they lie under /root/reference into oracle/_ref/ (git-ignored, travels to
# the GPU box via gpurun).  Nothing from the reference is copied into this repo.
# Flags mirror the reference's Release build (CMakeLists.txt:88-89,109,134-137):
# -O3 -DNDEBUG + OpenMP.  The reference's own CMake is NOT run (it needs
# BLAS/LAPACK/nifticlib/DCMTK discovery that fails here).
#
# LAPACK (dsyevd_ for eigen_Mat_rm, imutil.c:3035) comes from the OpenBLAS 0.3.15
# bundled with the image's opencv wheel; if that file is absent the script falls
# back to oracle/lapack_shim.c (a Jacobi dsyevd_ written for the oracle only).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${SIFT3D_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$R/imutil" ]; then
  echo "build_ref: $R not present (GPU box?) - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
SITE="$(python - <<'PY'
import sysconfig; print(sysconfig.get_paths()["purelib"])
PY
)"
OBD="$SITE/opencv_python_headless.libs"
BLAS="$(ls "$OBD"/libopenblasp-*.so 2>/dev/null | head -1 || true)"
DEFS="-DSIFT3D_VERSION_NUMBER=1.4.6 -DSIFT3D_HAVE_STRNLEN -DSIFT3D_HAVE_STRNDUP"
CF="-O3 -DNDEBUG -fopenmp -fPIC -w"
if [ -n "$BLAS" ]; then
  LAPACK_LINK="-Wl,--no-as-needed $BLAS -Wl,--disable-new-dtags -Wl,-rpath,$OBD"
  echo "openblas:$BLAS" > "$OUT/LAPACK_PROVIDER"
else
  gcc -O2 -fPIC -c "$HERE/lapack_shim.c" -o "$OUT/lapack_shim.o"
  LAPACK_LINK="$OUT/lapack_shim.o"
  echo "shim:oracle/lapack_shim.c" > "$OUT/LAPACK_PROVIDER"
fi
gcc $CF $DEFS -I"$R/imutil" -shared "$R/imutil/imutil.c" "$R/imutil/nifti.c" \
    -x c++ "$R/imutil/dicom.cpp" -x none -o "$OUT/libimutil_ref.so" \
    $LAPACK_LINK -lz -lm -lstdc++
gcc $CF $DEFS -I"$R/imutil" -I"$R/sift3d" -shared "$R/sift3d/sift.c" \
    -o "$OUT/libsift3D_ref.so" -L"$OUT" -limutil_ref -lm -Wl,-rpath,'$ORIGIN'
# The reference's UNMODIFIED command-line programs, for the end-to-end drop-in test
# (tests/test_gpu_cli.py, BASELINE.json configs[0]): *_stock resolves libsift3D.so at run time
# (the test points LD_LIBRARY_PATH at sift3d_b200/lib), *_cpu is bound to the reference build.
ln -sf libsift3D_ref.so "$OUT/libsift3D.so.cpu"
for prog in kpSift3D denseSift3D; do
  gcc -O1 -w $DEFS -I"$R/imutil" -I"$R/sift3d" "$R/cli/$prog.c" -o "$OUT/${prog}_stock" \
      -L"$HERE/../sift3d_b200/lib" -lsift3D -L"$OUT" -limutil_ref -lm \
      -Wl,-rpath,'$ORIGIN' -Wl,--allow-shlib-undefined 2>/dev/null || \
      echo "build_ref: ${prog}_stock not linked (build sift3d_b200/lib first)" >&2
  gcc -O1 -w $DEFS -I"$R/imutil" -I"$R/sift3d" "$R/cli/$prog.c" -o "$OUT/${prog}_cpu" \
      -L"$OUT" -lsift3D_ref -limutil_ref -lm -Wl,-rpath,'$ORIGIN' -Wl,--allow-shlib-undefined
done
# the reference's example volume (data, not source; _ref/ is git-ignored and travels to the box)
mkdir -p "$OUT/data"
cp -f "$R/examples/data/1.nii.gz" "$OUT/data/1.nii.gz"
echo "build_ref: built $(ls "$OUT")"
