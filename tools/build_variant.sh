#!/bin/bash
# A/B helper: tools/build_variant.sh NAME "EXTRA_NVCC_FLAGS" obj1 obj2 ...   (obj = inst_vdp_mu, impl_strict_robertson, ...)
# Compiles the listed objects with the extra flags into build/var_NAME and links ivp_b200/lib/libivpb_NAME.so from them
# plus the untouched objects of the main build.  Use with IVPB_LIB=ivp_b200/lib/libivpb_NAME.so python bench.py ...
set -e
cd "$(dirname "$0")/../ivp_b200/csrc"
NAME=$1; EXTRA=$2; shift 2
MAIN=../../build/ivpb; VAR=../../build/var_$NAME
mkdir -p $VAR
NV="nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -diag-suppress 177 -Xfatbin -compress-all $EXTRA"
struct_of() { grep -o "$1:[A-Za-z0-9]*" Makefile | head -1 | cut -d: -f2; }
REPL=""
for o in "$@"; do
  (
  case $o in
    inst_strictd_*) t=${o#inst_strictd_}; $NV -fmad=false -DIVPB_STRICT -DIVPB_DEFER_GUARDS -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst.cu -o $VAR/$o.o ;;
    impl_strictd_*) t=${o#impl_strictd_}; $NV -fmad=false -DIVPB_STRICT -DIVPB_DEFER_GUARDS -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst_implicit.cu -o $VAR/$o.o ;;
    inst_strict_*) t=${o#inst_strict_}; $NV -fmad=false -DIVPB_STRICT -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst.cu -o $VAR/$o.o ;;
    inst_*) t=${o#inst_}; $NV -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst.cu -o $VAR/$o.o ;;
    impl_strict_*) t=${o#impl_strict_}; $NV -fmad=false -DIVPB_STRICT -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst_implicit.cu -o $VAR/$o.o ;;
    impl_*) t=${o#impl_}; $NV -DIVPB_PROBLEM=$(struct_of $t) -DIVPB_TAG=$t -c ivpb_inst_implicit.cu -o $VAR/$o.o ;;
  esac
  ) &
done
wait
OBJS=""
for f in $MAIN/*.o; do b=$(basename $f .o); if [ -f $VAR/$b.o ]; then OBJS="$OBJS $VAR/$b.o"; else OBJS="$OBJS $f"; fi; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../lib/libivpb_$NAME.so $OBJS -lcudart_static -lpthread -ldl -lrt
ls -la ../lib/libivpb_$NAME.so
