# build a variant of the library with extra -D definitions for the fused similarity kernel only:
#   bash tools/build_variant.sh libovdet_x1.so -DOVDET_X_NOTMA
# (the in-tree libovdet.so must have been built; the other objects are reused) - for tools/ab_lib.sh
OUT=$1; shift
P=$(ls -d real-time-*_b200)
nvcc "$@" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden \
  -I include -c $P/csrc/sim_fused_sm100.cu -o /tmp/sim_fused_variant_$$.o || exit 1
OBJS=$(ls $P/csrc/_obj/*.o | grep -v sim_fused_sm100.o)
nvcc -shared -o $P/$OUT $OBJS /tmp/sim_fused_variant_$$.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC && echo built $P/$OUT
rm -f /tmp/sim_fused_variant_$$.o
