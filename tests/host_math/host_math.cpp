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

// ---- the block-level code of the packed kernels (csrc/svs_block.cuh) on the host --------------
// One 8x8 block per call: `px` = 8 rows x 8*ch bytes, `bits` = n payload bits (0/1 bytes).
#include "svs_block.cuh"

namespace {
void words_of(const uint8_t* px, int ch, uint32_t* rows)
{
    const int P = ch == 3 ? 6 : 2;
    for (int r = 0; r < 8; ++r)
        for (int k = 0; k < P; ++k) {
            uint32_t w = 0;
            for (int b = 0; b < 4; ++b) w |= (uint32_t)px[r * 8 * ch + 4 * k + b] << (8 * b);
            rows[r * P + k] = w;
        }
}
void bytes_of(const uint32_t* w16, uint8_t* out)
{
    for (int k = 0; k < 16; ++k)
        for (int b = 0; b < 4; ++b) out[4 * k + b] = (uint8_t)(w16[k] >> (8 * b));
}
}  // namespace

extern "C" {
// returns 0, or -1 when the packed quantiser does not cover this delta (the kernels then use the scalar path)
int hm_blk_embed(int ch, const uint8_t* px, long nblocks, double delta, int n, const uint8_t* bits, uint8_t* stego, uint8_t* gray)
{
    const svs::FastQuant fq = svs::make_fast_quant(delta);
    if (!fq.embed_ok || (double)(float)delta != delta) return -1;
    const blk::QuantRegs Q = blk::make_quant_regs(fq, (float)delta);
    for (long b = 0; b < nblocks; ++b) {
        uint32_t rows[48], s[16], g[16], w0 = 0, w1 = 0;
        words_of(px + b * 64 * ch, ch, rows);
        for (int i = 0; i < n; ++i) {
            if (i < 32) w0 |= (uint32_t)(bits[b * n + i] & 1) << (31 - i);
            else w1 |= (uint32_t)(bits[b * n + i] & 1) << (63 - i);
        }
#define HM_E(CH, NF) blk::block_embed<CH, NF, true>(rows, 0x4B000000u, Q, n, w0, w1, s, g)
        if (ch == 3) { if (n == 63) HM_E(3, true); else HM_E(3, false); }
        else         { if (n == 63) HM_E(1, true); else HM_E(1, false); }
#undef HM_E
        bytes_of(s, stego + b * 64);
        bytes_of(g, gray + b * 64);
    }
    return 0;
}

int hm_blk_extract(int ch, const uint8_t* px, long nblocks, double delta, int n, uint8_t* bits_out)
{
    const svs::FastQuant fq = svs::make_fast_quant(delta);
    if (!fq.extract_ok) return -1;
    const blk::QuantRegs Q = blk::make_quant_regs(fq, (float)delta);
    const int np = (n + 1 + 15) / 16;
    for (long b = 0; b < nblocks; ++b) {
        uint32_t rows[48], hi = 0, lo = 0;
        words_of(px + b * 64 * ch, ch, rows);
#define HM_X(CH, NP) blk::block_extract<CH, NP>(rows, 0x4B000000u, Q, n, hi, lo)
        if (ch == 3) { if (np == 1) HM_X(3, 1); else if (np == 2) HM_X(3, 2); else if (np == 3) HM_X(3, 3); else HM_X(3, 4); }
        else         { if (np == 1) HM_X(1, 1); else if (np == 2) HM_X(1, 2); else if (np == 3) HM_X(1, 3); else HM_X(1, 4); }
#undef HM_X
        for (int i = 0; i < n; ++i) bits_out[b * n + i] = (uint8_t)(i < 32 ? (hi >> (31 - i)) & 1u : (lo >> (63 - i)) & 1u);
        // nothing may be set beyond bit n
        const unsigned long long s = ((unsigned long long)hi << 32) | lo;
        if (n < 64 && (s << n) != 0) return -2;
    }
    return 0;
}
}
