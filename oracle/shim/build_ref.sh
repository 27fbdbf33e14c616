#!/bin/sh
# Compiles the reference's own source files, UNMODIFIED and from where they lie, against the header shims
# in this directory (Eigen / OpenCV / CppAD are not installed in this image).  Outputs go to oracle/_ref/
# only (git-ignored; they travel to the GPU box with the snapshot).  Usage: build_ref.sh /root/reference
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT="$HERE/../_ref"
mkdir -p "$OUT"
# shim dir first (so "cubic_spline.h" resolves to the stand-in), then the reference's own include/
FLAGS="-std=c++11 -O2 -ffp-contract=off -fPIC -shared -w -I$HERE -I$REF/include"
g++ $FLAGS -DWHICH=1 -DREF_SRC="\"$REF/src/extended_kalman_filter.cpp\"" "$HERE/ref_wrap.cpp" -o "$OUT/libref_ekf.so"
g++ $FLAGS -DWHICH=2 -DREF_SRC="\"$REF/src/particle_filter.cpp\"" "$HERE/ref_wrap.cpp" -o "$OUT/libref_pf.so"
g++ $FLAGS -DWHICH=3 -DREF_SRC="\"$REF/src/model_predictive_control.cpp\"" "$HERE/ref_wrap.cpp" -o "$OUT/libref_mpc.so"
g++ $FLAGS -DWHICH=4 -DREF_SRC="\"$REF/src/lqr_steer_control.cpp\"" "$HERE/ref_wrap.cpp" -o "$OUT/libref_lqr4.so"
g++ $FLAGS -DWHICH=5 -DREF_SRC="\"$REF/src/lqr_speed_steer_control.cpp\"" "$HERE/ref_wrap.cpp" -o "$OUT/libref_lqr5.so"
echo "oracle/_ref: libref_lqr4.so libref_lqr5.so libref_ekf.so libref_pf.so libref_mpc.so built from $REF/src"
