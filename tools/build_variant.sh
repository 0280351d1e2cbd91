#!/bin/bash
# Builds a second copy of the C-ABI library with extra nvcc flags (kernel A/B experiments; select it with AE_LIB=<path>).
#   tools/build_variant.sh red4 -DAE_SWEEP_REDUCERS=4   ->  alpha_eigen_b200/libae_red4.so
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/alpha_eigen_b200/csrc
out=$root/build/variant_$name
mkdir -p "$out"
for f in ae_core ae_sweep ae_small ae_source ae_solver ae_dmd ae_comm; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -I"$root/include" "$@" \
      -c "$src/$f.cu" -o "$out/$f.o" &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "$root/alpha_eigen_b200/libae_$name.so" "$out"/*.o \
    -L/usr/local/cuda/lib64 -lcusolver -lcublas -ldl -Xlinker -rpath,/usr/local/cuda/lib64
echo "built alpha_eigen_b200/libae_$name.so"
