#!/bin/bash
# Build a kernel-variant library (development aid): scripts/build_variant.sh <suffix> "<-D flags>"
# Objects that do not depend on the filter-kernel knobs are reused from the default build.
cd "$(dirname "$0")/../alfred-margaret_b200"
sfx=$1; shift
mkdir -p build$sfx
for o in am_kernels am_synth am_replacer; do [ -f build/$o.o ] && cp -p build/$o.o build$sfx/ && touch build$sfx/$o.o; done
make -j4 OUT=lib/libam_b200$sfx.so BUILD=build$sfx VARIANT="$*" 2>&1 | grep -E "error|warning" 
ls -la lib/libam_b200$sfx.so
