// Host-only build of csrc/svs_math.cuh (the exact operation sequence the CUDA kernels run),
// so the CPU test-suite can compare it with the oracle bit for bit.  Test infrastructure only.
// Compile with: g++ -O2 -std=c++17 -ffp-contract=off -fno-fast-math -shared -fPIC
#include "svs_math.cuh"
#include "svs_quant.h"

extern "C" {
void hm_dct2_fwd(float* blocks, long n) { for (long i = 0; i < n; ++i) svs::dct2_fwd(svs::ScalarOps(), blocks + 64 * i); }
void hm_dct2_inv(float* blocks, long n) { for (long i = 0; i < n; ++i) svs::dct2_inv(svs::ScalarOps(), blocks + 64 * i); }
void hm_dct8_fwd(float* rows, long n) { for (long i = 0; i < n; ++i) svs::dct8_fwd<1>(svs::ScalarOps(), rows + 8 * i); }
void hm_dct8_inv(float* rows, long n) { for (long i = 0; i < n; ++i) svs::dct8_inv<1>(svs::ScalarOps(), rows + 8 * i); }

// Quantiser claims (svs_quant.h): for every coefficient c and both payload bits compare the
// speculative and the exact quantiser with the reference's IEEE division.  stats[0..7] =
// flagged embed, flagged extract, unflagged-but-different embed, ... extract, exact-path
// different embed, ... extract, embed_ok, extract_ok.
void hm_quant_check(double delta, const float* c, long n, long* stats)
{
    const svs::FastQuant q = svs::make_fast_quant(delta);
    const float d = (float)delta;
    for (int i = 0; i < 8; ++i) stats[i] = 0;
    stats[6] = q.embed_ok;
    stats[7] = q.extract_ok;
    for (long i = 0; i < n; ++i) {
        bool flag;
        if (q.embed_ok)
            for (int bit = 0; bit < 2; ++bit) {
                const float want = svs::ref_embed(c[i], d, bit);
                const float got = svs::fast_embed(q, c[i], bit, flag);
                if (bit == 0) stats[0] += flag;
                if (!flag && svs::f2u(got) != svs::f2u(want)) stats[2]++;
                if (svs::f2u(svs::exact_embed(c[i], d, q.r, bit)) != svs::f2u(want)) stats[4]++;
            }
        if (q.extract_ok) {
            const int want = svs::ref_parity(c[i], d);
            const int got = svs::fast_parity(q, c[i], flag);
            stats[1] += flag;
            if (!flag && got != want) stats[3]++;
            if (svs::exact_parity(c[i], d, q.r) != want) stats[5]++;
        }
    }
}
}
