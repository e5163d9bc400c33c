#!/bin/bash
# tests/integration/build.sh — compiles the reference-side binding (gpu_process.cpp) against the REFERENCE's own sources
# and links it with libsnk_engine.so:  oracle/_ref/SOAPnuke_gpu  (git-ignored, travels to the GPU box).
#
# The reference tree is read-only and none of it is copied into this repository: the sources are copied to a scratch
# directory, patched there, compiled, and the scratch directory is removed. The whole patch a maintainer would apply:
#
#   src/peprocess.h:60   void *stat_pe_fqs(PEstatOption opt, string dataType);   ->  virtual void *stat_pe_fqs(...)
#   src/peprocess.h:69   void merge_stat();                                      ->  virtual void merge_stat();
#   src/seprocess.h:38   void* stat_se_fqs(SEstatOption opt,string dataType);    ->  virtual void* stat_se_fqs(...)
#   src/seprocess.h:40   void filter_se_fqs(SEcalOption opt);                    ->  virtual void filter_se_fqs(...)
#   src/seprocess.h:49   void merge_stat();                                      ->  virtual void merge_stat();
#   src/main.cpp:58,62   peProcess new_task(gp); / seProcess new_task(gp);       ->  gpuPeProcess / gpuSeProcess (+ #include "gpu_process.h")
#   Makefile             add gpu_process.cpp, <repo>/soapnuke_b200/host/{cli_params,host_common}.cpp, -lsnk_engine
#
# filter_pe_fqs is already virtual (peprocess.h:61). Nothing else of the reference changes: its reader threads, temp
# files, `cat`, update_stat and print_stat run as they are.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${REF:-/root/reference}
OUT=$ROOT/oracle/_ref/SOAPnuke_gpu
if [ ! -d "$REF/src" ]; then echo "reference sources not present ($REF); using prebuilt $OUT if any"; exit 0; fi
if [ -x "$OUT" ] && [ "$OUT" -nt "$HERE/gpu_process.cpp" ] && [ "$OUT" -nt "$HERE/gpu_process.h" ] && [ "$OUT" -nt "$ROOT/include/snk_engine.h" ] \
   && [ "$OUT" -nt "$ROOT/soapnuke_b200/host/cli_params.cpp" ] && [ "$OUT" -nt "$HERE/build.sh" ]; then exit 0; fi
W=$(mktemp -d /tmp/snk_integration.XXXXXX)
trap 'rm -rf "$W"' EXIT
cp "$REF"/src/*.cpp "$REF"/src/*.h "$W"/
rm -f "$W"/mGzip.cpp "$W"/processHts.cpp
sed -i 's/^\tvoid \*stat_pe_fqs(/\tvirtual void *stat_pe_fqs(/; s/^\tvoid merge_stat();/\tvirtual void merge_stat();/' "$W"/peprocess.h
sed -i 's/^\tvoid\* stat_se_fqs(/\tvirtual void* stat_se_fqs(/; s/^\tvoid filter_se_fqs(SEcalOption opt);/\tvirtual void filter_se_fqs(SEcalOption opt);/; s/^\tvoid merge_stat();/\tvirtual void merge_stat();/' "$W"/seprocess.h
sed -i 's/peProcess new_task(gp);/gpuPeProcess new_task(gp);/; s/seProcess new_task(gp);/gpuSeProcess new_task(gp);/; s/#include "seprocess.h"/#include "seprocess.h"\n#include "gpu_process.h"/' "$W"/main.cpp
grep -q "virtual void \*stat_pe_fqs" "$W"/peprocess.h && grep -q "virtual void merge_stat" "$W"/peprocess.h
grep -q "virtual void\* stat_se_fqs" "$W"/seprocess.h && grep -q "virtual void filter_se_fqs" "$W"/seprocess.h && grep -q "virtual void merge_stat" "$W"/seprocess.h
grep -q "gpuPeProcess new_task" "$W"/main.cpp && grep -q "gpuSeProcess new_task" "$W"/main.cpp
mkdir -p "$ROOT/oracle/_ref"
echo "building the reference with the engine binding -> $OUT"
# the reference's sources need `-include cstdint` with gcc 13 (see oracle/Makefile); the binding and the host helpers are C++17
for f in "$W"/*.cpp; do g++ -std=c++11 -O2 -w -include cstdint -I"$W" -I"$HERE" -I"$ROOT/include" -c "$f" -o "${f%.cpp}.o" & done; wait
g++ -std=c++17 -O2 -w -include cstdint -I"$W" -I"$ROOT/include" -I"$ROOT/soapnuke_b200/host" -I"$ROOT/soapnuke_b200/csrc" \
    -c "$HERE/gpu_process.cpp" -o "$W/gpu_process.o"
g++ -std=c++17 -O2 -w -c "$ROOT/soapnuke_b200/host/cli_params.cpp" -o "$W/snk_cli_params.o"
g++ -std=c++17 -O2 -w -c "$ROOT/soapnuke_b200/host/host_common.cpp" -o "$W/snk_host_common.o"
g++ -o "$OUT" "$W"/*.o -L"$ROOT/soapnuke_b200/lib" -lsnk_engine -lz -lpthread -Wl,-rpath,'$ORIGIN/../../soapnuke_b200/lib'
ls -la "$OUT"
