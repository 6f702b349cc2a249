#!/bin/bash
# Developer: build liblgae_b200 with extra -D flags into tools/variants/<name>.so (git-ignored; travels with gpurun).
# usage: tools/build_variant.sh <name> -DLGAE_RPD_BWD=4 ...      then   LGAE_B200_LIB=tools/variants/<name>.so python bench.py
name=$1; shift
src=lgn_autoencoder_b200/csrc
objs=""
mkdir -p /tmp/variant_$name
for f in lgae_api lgae_glue lgae_level lgae_radial lgae_mlp lgae_cg lgae_layers lgae_optim lgae_collective; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c $src/$f.cu -o /tmp/variant_$name/$f.o &
  objs="$objs /tmp/variant_$name/$f.o"
done
wait
nvcc -shared -o tools/variants/$name.so $objs -gencode arch=compute_100a,code=sm_100a && echo built tools/variants/$name.so
