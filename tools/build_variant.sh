#!/bin/bash
# One variant build of the library: tools/build_variant.sh <name> <file.cu under csrc, or a path to an alternative source for it> <basename.cu it replaces> [nvcc -D flags...]
# -> tools/bin/libhulk_<name>.so (every other object from the current in-tree build); select it with HK_LIB_PATH.
set -e
cd "$(dirname "$0")/.."
name=$1; src=$2; repl=$3; shift 3
python -m hulk_keypoints_b200.build > /dev/null
B=hulk_keypoints_b200/csrc/build
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
OTHERS=$(ls $B/*.o | grep -v "\.diag\.o" | grep -v "/${repl%.cu}\.o")
mkdir -p tools/bin
nvcc $FLAGS "$@" -Ihulk_keypoints_b200/csrc -c $src -o /tmp/variant_$name.o
nvcc -shared -o tools/bin/libhulk_$name.so /tmp/variant_$name.o $OTHERS -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -lcudart
echo built tools/bin/libhulk_$name.so
