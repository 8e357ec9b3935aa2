#!/bin/bash
# developer tool: build a variant of the library with extra -D flags into build_variants/<name>/libmcpower_b200.so
# (only the translation units named in $TUS are recompiled with the flags; the rest is linked from the main object cache).
# usage: TUS=1 tools/build_variant.sh <name> -DMCP_LEAN_SCAN=1 ...     (safe to run several at once)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
obj=/tmp/mcpower_b200_build_$name
mkdir -p "$obj" "$root/build_variants/$name"
cp -u /tmp/mcpower_b200_build/*.o "$obj"/
cd "$root"
MCPOWER_B200_OBJ_DIR=$obj MCPOWER_B200_SO_OUT=$root/build_variants/$name/libmcpower_b200.so python -m mcpower_b200.build --tus ${TUS:-1} "$@" > "$obj/build.log" 2>&1 || { tail -20 "$obj/build.log"; exit 1; }
echo "built build_variants/$name"
