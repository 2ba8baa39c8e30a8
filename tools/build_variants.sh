#!/bin/bash
# Builds the prepared but not yet measured kernel variants into bifrost3d_b200/variants/ (run HERE, then
# `gpurun -- tools/run_variants.sh`). Each variant is the default build plus one macro.
set -e
cd "$(dirname "$0")/../bifrost3d_b200/csrc"
mkdir -p ../variants
build() { make -s BUILD=build_$1 OUT=../variants/libbpt_$1.so EXTRA="$2" -j8 && echo "built $1 ($2)"; }
build odiv   "-DBPT_OUTLINE_DIV=1"
build odsq   "-DBPT_OUTLINE_DIV=1 -DBPT_OUTLINE_SQRT=1"
build pl2    "-DBPT_PARKED_LEAVES=2"
build sb64   "-DBPT_SHADE_BLOCK=64 -DBPT_SHADE_MIN_BLOCKS=16"
build odsqpl "-DBPT_OUTLINE_DIV=1 -DBPT_OUTLINE_SQRT=1 -DBPT_PARKED_LEAVES=2"
build split  "-DBPT_SHADE_SPLIT=1"
build splito "-DBPT_SHADE_SPLIT=1 -DBPT_OUTLINE_DIV=1 -DBPT_OUTLINE_SQRT=1"
