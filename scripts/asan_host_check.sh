#!/bin/bash
# AddressSanitizer pass over the HOST logic of libsipgpu (no GPU needed): a scratch build of aces4_b200/csrc with
# -fsanitize=address in which the two monotonic arenas of worklist.cu have no retained buffer (so release() really frees and a
# use-after-release is a reportable use-after-free), driven by the dry-mode pardo stream and the CPU host-logic tests.
#   bash scripts/asan_host_check.sh            -> profiles/r01_asan_host.txt
set -eu
ROOT=$(cd "$(dirname "$0")/.." && pwd)
W=$(mktemp -d)
mkdir -p "$W/a/b" "$W/lib"
cp -r "$ROOT/aces4_b200/csrc" "$W/a/b/csrc"; rm -rf "$W/a/b/csrc/build"; cp -r "$ROOT/include" "$W/a/include"
python - "$W/a/b/csrc/worklist.cu" <<'PY'
import sys
p = sys.argv[1]; s = open(p).read()
a = s.index('std::pmr::monotonic_buffer_resource& rec_mem() {'); b = s.index('using PairVec')
s = s[:a] + ('std::pmr::monotonic_buffer_resource& rec_mem() { static std::pmr::monotonic_buffer_resource r; return r; }\n'
             'std::pmr::monotonic_buffer_resource& sched_mem() { static std::pmr::monotonic_buffer_resource r; return r; }\n') + s[b:]
open(p, 'w').write(s)
PY
( cd "$W/a/b/csrc" && make -j8 OUT="$W/lib/libsipgpu.so" EXTRA="-Xcompiler -fsanitize=address,-fno-omit-frame-pointer -g" >/dev/null 2>&1 \
  || /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --version-script=exports.map \
       -Xcompiler -fsanitize=address -o "$W/lib/libsipgpu.so" build/*.o -lcudart )
OUT="$ROOT/profiles/r01_asan_host.txt"
{
  echo "# AddressSanitizer, host logic of libsipgpu (scripts/asan_host_check.sh); arenas of worklist.cu heap-backed so that release() frees"
  g++ -O1 -g -fsanitize=address -std=c++17 "$ROOT/scripts/micro/wl_dry_bench.cpp" -I"$ROOT/include" -L"$W/lib" -lsipgpu -Wl,-rpath,"$W/lib" -o "$W/wl_dry_asan"
  echo "## wl_dry_bench 16 3 16 6 (34 992 recorded ops)"; ASAN_OPTIONS=detect_leaks=0 "$W/wl_dry_asan" 16 3 16 6 2 2>&1 | tail -3
  echo "## pytest (CPU host-logic tests) with the sanitized library preloaded"
  cd "$ROOT" && LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 SIPGPU_LIB="$W/lib/libsipgpu.so" \
    python -m pytest tests/test_worklist_cpu.py tests/test_consistency_cpu.py tests/test_mirror_cpu.py tests/test_persist_cpu.py \
      tests/test_planner_cpu.py tests/test_sial_frontend_cpu.py -q \
      --deselect tests/test_worklist_cpu.py::test_lccd_pardo_stream_dry_recording_counts_and_host_cost 2>&1 | tail -3
  echo "## AddressSanitizer reports: $(grep -c 'ERROR: AddressSanitizer' "$OUT.tmp" 2>/dev/null || echo 0)"
} > "$OUT.tmp" 2>&1 || true
grep -c "ERROR: AddressSanitizer" "$OUT.tmp" > /dev/null && sed -i "s/^## AddressSanitizer reports:.*/## AddressSanitizer reports: $(grep -c 'ERROR: AddressSanitizer' "$OUT.tmp")/" "$OUT.tmp" || true
mv "$OUT.tmp" "$OUT"; rm -rf "$W"; cat "$OUT"
