#!/bin/bash
# Debug / experiment build of the library next to the product one:  bash tools/build_variant.sh timers -DBUCTD_ROLE_TIMERS
#   -> buctd_b200/libbuctd_<name>.so  (use with BUCTD_LIB=buctd_b200/libbuctd_<name>.so)
set -e
name=$1; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
bd=/tmp/buctd_build_$name
mkdir -p $bd/buctd_b200/csrc $bd/include
cp -u $root/buctd_b200/csrc/*.cu $root/buctd_b200/csrc/*.cuh $root/buctd_b200/csrc/*.h $root/buctd_b200/csrc/Makefile $bd/buctd_b200/csrc/
cp -u $root/include/*.h $bd/include/
make -s -C $bd/buctd_b200/csrc -j 8 EXTRA="$*" OUT=$root/buctd_b200/libbuctd_$name.so
ls -la $root/buctd_b200/libbuctd_$name.so
