#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never on the product path.
#
# Compiles the UNMODIFIED reference (ad3002/aindex, /root/reference/src) with the
# reference's own flags (Makefile:3: g++ -std=c++17 -pthread -O3 -fPIC) from the
# sources where they lie; outputs go only into oracle/_ref/ (git-ignored, but it
# travels to the GPU box with the snapshot).  No reference source is copied.
#
# Products:
#   oracle/_ref/bin/{compute_mphf_seq,compute_index,compute_aindex,compute_reads,
#                    count_kmers13,compute_aindex13,generate_all_13mers}
#   oracle/_ref/aindex_cpp*.so      the reference pybind11 module (import directly)
#   oracle/_ref/bin/ref_harness     oracle/ref_harness.cpp built against the
#                                   reference's hash.hpp/kmers.hpp (threaded get_freq)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${AINDEX_REFERENCE:-/root/reference}/src"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT/obj" "$OUT/bin"
CXX="${ORC_CXX:-/usr/bin/g++}"  # not $CXX: the image's /opt/gcc wrapper links libstdc++ statically, which clashes with numpy's in one process
FLAGS="-std=c++17 -pthread -O3 -fPIC -w"
for f in helpers debrujin read kmers settings hash; do
  if [ ! -f "$OUT/obj/$f.o" ]; then $CXX $FLAGS -I"$REF" -c "$REF/$f.cpp" -o "$OUT/obj/$f.o" & fi
done
wait
OBJS="$OUT/obj/helpers.o $OUT/obj/debrujin.o $OUT/obj/read.o $OUT/obj/kmers.o $OUT/obj/settings.o $OUT/obj/hash.o"
[ -f "$OUT/bin/compute_mphf_seq" ] || $CXX $FLAGS -I"$REF" "$REF/emphf/compute_mphf_seq.cpp" -o "$OUT/bin/compute_mphf_seq" &
for t in compute_index compute_aindex compute_reads count_kmers13 compute_aindex13; do
  [ -f "$OUT/bin/$t" ] || $CXX $FLAGS -I"$REF" "$REF/$t.cpp" $OBJS -o "$OUT/bin/$t" &
done
[ -f "$OUT/bin/generate_all_13mers" ] || $CXX $FLAGS -I"$REF" "$REF/generate_all_13mers.cpp" "$OUT/obj/kmers.o" -o "$OUT/bin/generate_all_13mers" &
wait
PYINC="$(python3 -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
PBINC="$(python3 -c 'import pybind11;print(pybind11.get_include())')"
EXT="$(python3 -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
[ -f "$OUT/aindex_cpp$EXT" ] || $CXX $FLAGS -shared -I"$PYINC" -I"$PBINC" -I"$REF" "$REF/python_wrapper.cpp" $OBJS -o "$OUT/aindex_cpp$EXT" &
if [ -f "$HERE/ref_harness.cpp" ]; then
  if [ ! -f "$OUT/bin/ref_harness" ] || [ "$HERE/ref_harness.cpp" -nt "$OUT/bin/ref_harness" ]; then
    $CXX $FLAGS -I"$REF" "$HERE/ref_harness.cpp" $OBJS -o "$OUT/bin/ref_harness" &
  fi
fi
wait
echo "build_ref: ok -> $OUT"
