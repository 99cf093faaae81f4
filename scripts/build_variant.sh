# Build an A/B variant of the CUDA library:  bash scripts/build_variant.sh <name> [-DFLAG=... ...]
# -> fountain_b200/csrc/libfountain_gpu_<name>.so  (select it at run time with FTN_GPU_LIB)
set -e
NAME=$1; shift
cd "$(dirname "$0")/../fountain_b200/csrc"
mkdir -p build_$NAME
for f in capi scene sort_scan trace render; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-fvisibility=hidden --expt-relaxed-constexpr "$@" -c $f.cu -o build_$NAME/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o libfountain_gpu_$NAME.so build_$NAME/*.o -lcudart
rm -rf build_$NAME
echo built libfountain_gpu_$NAME.so
