#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (biointec/blamm) from the sources where
# they lie under /root/reference into the git-ignored oracle/_ref/.  Nothing here is product code and no
# reference source is copied into the repository.  Recipe: SURVEY.md section 8c.
#   * cmake/FortranScheme.cmake would emit  F77_FUNC(name,NAME) name##_  for gfortran -> hand-written config.h
#   * species.h uses std::array without <array>  ->  -include array
#   * BLAS = OpenBLAS 0.3.15 that ships inside the opencv wheel of this image (exports plain sgemm_)
# Outputs:  oracle/_ref/blamm          the reference CLI (CPU BLAS path + naive path)
#           oracle/_ref/refdump        float-level harness linking the reference's own classes (refdump.cpp)
#           oracle/_ref/blamm_cuda     the reference CLI with its own GPU path (`scan -c`: cuBLAS sgemm + kernel.cu's filterScore,
#                                      pwmscan.cpp:297-437) compiled for sm_100 -- the "existing GPU implementation" bench.py times
#                                      beside the B200-native path; only built where nvcc exists
#           oracle/_ref/blamm_dropin   the reference CLI with INTEGRATION.md's three edits + integration/scanPWMB200.inc, linked against
#                                      libb200scan.so: its `scan -c` runs the B200-native path (build_dropin.py; tests/test_reference_dropin.py)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${BLAMM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OB="${OPENBLAS_DIR:-/opt/prime-rl/.venv/lib/python3.12/site-packages/opencv_python_headless.libs}"
OBLIB="$(ls "$OB" | grep -m1 '^libopenblas' || true)"
if [ ! -d "$REF/src" ]; then echo "build_ref: $REF/src absent (GPU box?) - keeping prebuilt files"; exit 0; fi
if [ -z "$OBLIB" ]; then echo "build_ref: no OpenBLAS under $OB"; exit 1; fi
mkdir -p "$OUT/cfg"
cat > "$OUT/cfg/config.h" <<'EOC'
#define F77_FUNC(name,NAME) name ## _
EOC
CXXFLAGS="-O3 -std=c++11 -DNDEBUG -DHAVE_CONFIG_H -DBLAMM_MAJOR_VERSION=1 -DBLAMM_MINOR_VERSION=0 -DBLAMM_PATCH_LEVEL=0 -I$OUT/cfg -include array -w"
LDFLAGS="-L$OB -l:$OBLIB -Wl,--disable-new-dtags -Wl,-rpath,$OB -lpthread"
/usr/bin/g++ $CXXFLAGS "$REF"/src/*.cpp -o "$OUT/blamm" $LDFLAGS
/usr/bin/g++ $CXXFLAGS -I"$REF/src" "$HERE/refdump.cpp" "$REF"/src/{motif,sequence,species,settings,matrix}.cpp -o "$OUT/refdump" $LDFLAGS
if command -v nvcc > /dev/null 2>&1; then
  CUDA="$(dirname "$(dirname "$(command -v nvcc)")")"
  [ -d "$CUDA/include" ] || CUDA=/usr/local/cuda
  nvcc -O3 -gencode arch=compute_100,code=sm_100 -c "$REF/src/kernel.cu" -o "$OUT/cfg/kernel.o"
  /usr/bin/g++ $CXXFLAGS -DHAVE_CUDA -I"$CUDA/include" "$REF"/src/*.cpp "$OUT/cfg/kernel.o" -o "$OUT/blamm_cuda" $LDFLAGS \
      -L"$CUDA/lib64" -Wl,-rpath,"$CUDA/lib64" -lcublas -lcudart
  echo "build_ref: built $OUT/blamm_cuda (reference GPU path, sm_100)"
fi
echo "$OB" > "$OUT/openblas_dir.txt"
echo "build_ref: built $OUT/blamm and $OUT/refdump"
# the reference with INTEGRATION.md's binding applied (temporary copy of its sources, removed again): `scan -c` goes through libb200scan.so
python3 "$HERE/build_dropin.py" "$REF" "$OUT" "$OB" "$OBLIB" || echo "build_ref: blamm_dropin not built"
