#!/bin/bash
# Builds A/B variants of librho_b200.so (kernel switches as -D flags) into rho_tts_b200/variants/ so that
# one gpurun call can time them side by side:   tools/ab_fused.sh name1 "-DFZ_BULK=0" name2 "-DFZ_FAST_APPLY=0" ...
# then on the GPU box:  python tools/ab_time.py
set -e
cd "$(dirname "$0")/../rho_tts_b200/csrc"
mkdir -p ../variants
rm -f ../variants/*.so
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC"
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  b=build_$name; mkdir -p $b
  for f in api host_api exchange join resample logmel cosine fused mel_gemm stft_tc qwen pitch mfcc spk; do $NV $flags -c $f.cu -o $b/$f.o & done
  $NV $flags -x cu -c tables.cpp -o $b/tables.o &
  g++ -O3 -std=c++17 -fPIC -c hostfill.cpp -o $b/hostfill.o &
  wait
  $NV -shared -o ../variants/lib_$name.so $b/*.o -cudart static -ldl
  rm -rf $b
  echo "built variants/lib_$name.so  [$flags]"
done
