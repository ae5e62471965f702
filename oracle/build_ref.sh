#!/usr/bin/env bash
# TEST INFRASTRUCTURE -- not part of the product path.
#
# Builds the reference's own C++ CEUpdater (cpp/src/*.cpp + the Cython shim
# cemc/cpp_ext/pyce_updater.pyx) from the sources WHERE THEY LIE under
# $CEMC_REFERENCE (default /root/reference) into oracle/_ref/ (git-ignored).
# Nothing from the reference is copied into the tracked tree; the only
# intermediate is a sed-patched temp copy of ce_updater.cpp (4 NumPy-2
# `(PyArrayObject*)` casts), deleted after compilation.
#
# We do not run the reference's setup.py: it builds ~20 unrelated translation
# units (Eshelby, Khachaturyan, Wang-Landau ...) that are out of scope.
set -euo pipefail
REF="${CEMC_REFERENCE:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF/cpp/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref (if any)"; exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
PY="${PYTHON:-python3}"
PYINC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
NPINC=$($PY -c "import numpy;print(numpy.get_include())")
SUFFIX=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
FLAGS="-std=c++11 -fopenmp -O3 -fPIC -w -DNDEBUG -DNPY_OUT_ARRAY=NPY_ARRAY_OUT_ARRAY -I$REF/cpp/include -I$PYINC -I$NPINC"
for f in cf_history_tracker additional_tools named_array row_sparse_struct_matrix cluster \
         symbols_with_numbers basis_function linear_vib_correction mc_observers init_numpy_api; do
  g++ $FLAGS -c "$REF/cpp/src/$f.cpp" -o "$TMP/$f.o" &
done
sed -e 's/PyArray_SIZE(npy_array)/PyArray_SIZE((PyArrayObject*)npy_array)/' \
    -e 's/PyArray_GETPTR1(npy_array,i)/PyArray_GETPTR1((PyArrayObject*)npy_array,i)/' \
    -e 's/PyArray_DIMS( trans_mat )/PyArray_DIMS((PyArrayObject*)trans_mat)/' \
    -e 's/PyArray_GETPTR2(trans_mat, i, col)/PyArray_GETPTR2((PyArrayObject*)trans_mat, i, col)/' \
    "$REF/cpp/src/ce_updater.cpp" > "$TMP/ce_updater_np2.cpp"
g++ $FLAGS -c "$TMP/ce_updater_np2.cpp" -o "$TMP/ce_updater.o" &
printf '# distutils: language = c++\n# cython: c_string_type=str, c_string_encoding=ascii\ninclude "%s/cemc/cpp_ext/pyce_updater.pyx"\n' "$REF" > "$TMP/cemc_cpp_code.pyx"
$PY -m cython --cplus -3 -I "$REF" -I "$REF/cemc/cpp_ext" "$TMP/cemc_cpp_code.pyx" -o "$TMP/cemc_cpp_code.cpp"
g++ $FLAGS -c "$TMP/cemc_cpp_code.cpp" -o "$TMP/cemc_cpp_code.o" &
wait
g++ -shared -fopenmp -o "$OUT/cemc_cpp_code$SUFFIX" "$TMP"/*.o
echo "build_ref: built $OUT/cemc_cpp_code$SUFFIX"
