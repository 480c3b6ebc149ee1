// Host-only build of csrc/svs_math.cuh (the exact operation sequence the CUDA kernels run),
// so the CPU test-suite can compare it with the oracle bit for bit.  Test infrastructure only.
// Compile with: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC
#include "svs_math.cuh"

extern "C" {
void hm_dct2_fwd(float* blocks, long n) { for (long i = 0; i < n; ++i) svs::dct2_fwd(svs::ScalarOps(), blocks + 64 * i); }
void hm_dct2_inv(float* blocks, long n) { for (long i = 0; i < n; ++i) svs::dct2_inv(svs::ScalarOps(), blocks + 64 * i); }
void hm_dct8_fwd(float* rows, long n) { for (long i = 0; i < n; ++i) svs::dct8_fwd<1>(svs::ScalarOps(), rows + 8 * i); }
void hm_dct8_inv(float* rows, long n) { for (long i = 0; i < n; ++i) svs::dct8_inv<1>(svs::ScalarOps(), rows + 8 * i); }
}
