#!/bin/bash
# Build a variant of the library under tools/variants/<name>.so (developer A/B timing): tools/build_variant.sh name [-DFOO=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/.."
objs=""
for f in lgae_api lgae_glue lgae_level lgae_radial lgae_mlp lgae_cg lgae_layers lgae_optim; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC "$@" -c lgn_autoencoder_b200/csrc/$f.cu -o /tmp/var_${name}_$f.o &
  objs="$objs /tmp/var_${name}_$f.o"
done
wait
nvcc -shared -o tools/variants/$name.so $objs -gencode arch=compute_100a,code=sm_100a
echo tools/variants/$name.so
