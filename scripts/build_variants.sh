#!/bin/bash
# Build tuning variants of libb2ode.so into gpurun_out-independent paths (tfdiffeq_b200/variants/*.so) for A/B runs.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p $ROOT/tfdiffeq_b200/variants
build() {
  name=$1; shift
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda \
    -Xcompiler -fPIC,-O3 -shared -I$ROOT/include -I$ROOT/tfdiffeq_b200/csrc --threads 2 "$@" \
    -o $ROOT/tfdiffeq_b200/variants/libb2ode_$name.so $ROOT/tfdiffeq_b200/csrc/b2ode.cu $ROOT/tfdiffeq_b200/csrc/b2ode_fused.cu
}
build minb6 -DB2_MINB_FINALIZE=6 -DB2_MINB_STAGE=6 &
build unroll2 -DB2_UNROLL=2 &
wait
build minb6_unroll2 -DB2_MINB_FINALIZE=6 -DB2_MINB_STAGE=6 -DB2_UNROLL=2 &
build minb8 -DB2_MINB_FINALIZE=8 -DB2_MINB_STAGE=8 &
wait
ls -la $ROOT/tfdiffeq_b200/variants
