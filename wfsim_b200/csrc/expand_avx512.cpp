// AVX-512 expander of the compact record transport (transport.cuh), branch-free per record:
// the fill pattern is a masked move (samples below `length`), the blocks of the stream are dropped
// into place with expand-loads (one mask bit per 8-byte block), and whole records leave the
// cache-resident scratch area as aligned 64-byte streaming stores.  Same bytes as expand_impl.inc.
#include <algorithm>
#include <stdint.h>
#include <string.h>
#include <immintrin.h>

#include "compact_format.h"

namespace wfs {

void expand_records_avx512(const CompactHdr *hdr, const uint8_t *blocks, int64_t j0, int64_t j1,
                           uint8_t *dst_base, int16_t fill, int16_t dt) {
    constexpr int kBatch = 64, kRec = 244, VB = 64;
    alignas(64) uint8_t scratch[64 + kRec * kBatch + 64];
    size_t carry = 0;
    uint8_t *dst = dst_base + (size_t)j0 * kRec;
    const __m512i wfill = _mm512_set1_epi16(fill);
    const uint64_t dt_bits = (uint64_t)(uint16_t)dt << 32;
    for (int64_t j = j0; j < j1; j += kBatch) {
        const int n = (int)std::min<int64_t>(kBatch, j1 - j);
        uint8_t *w = scratch + carry;
        for (int r = 0; r < n; r++, w += kRec) {
            const CompactHdr h = hdr[j + r];
            const uint32_t length = compact_length(h);
            // time | length, dt, channel | pulse_length, record_i, baseline = 0
            const uint64_t q0 = (uint64_t)h.time,
                           q1 = (uint64_t)length | dt_bits | ((uint64_t)(uint16_t)h.channel << 48),
                           q2 = (uint64_t)(uint32_t)h.pulse_length | ((uint64_t)(uint16_t)h.record_i << 32);
            memcpy(w, &q0, 8);
            memcpy(w + 8, &q1, 8);
            memcpy(w + 16, &q2, 8);
            uint8_t *d = w + 24;
            // one bit per sample below `length`; h.mask has one bit per 8-byte block in the stream
            const unsigned __int128 ones = ~(unsigned __int128)0;
            const unsigned __int128 smask = ~(ones << length);
            const uint32_t halves = h.mask;
            const uint8_t *src = blocks + (size_t)kBlockBytes * h.boff;
#pragma GCC unroll 4
            for (int v = 0; v < 4; v++) {      // 32 samples = 8 blocks = 64 bytes each (the last one spills
                                               // 36 bytes into the next record's slot, written after this one)
                const __mmask32 km = (__mmask32)(uint32_t)(smask >> (32 * v));
                const __mmask8 kb = (__mmask8)(halves >> (8 * v));
                const __m512i base = _mm512_maskz_mov_epi16(km, wfill);
                _mm512_storeu_si512(d + 64 * v, _mm512_mask_expandloadu_epi64(base, kb, src));
                src += 8 * _mm_popcnt_u32((uint32_t)kb);
            }
        }
        size_t total = carry + (size_t)n * kRec;
        const uint8_t *src = scratch;
        size_t head = (VB - ((uintptr_t)dst & (VB - 1))) & (VB - 1);
        if (head > total) head = total;
        if (head) { memcpy(dst, src, head); dst += head; src += head; total -= head; }
        const size_t nvec = total / VB;
        for (size_t v = 0; v < nvec; v++)
            _mm512_stream_si512(reinterpret_cast<__m512i *>(dst + v * VB), _mm512_loadu_si512(src + v * VB));
        dst += nvec * VB; src += nvec * VB; total -= nvec * VB;
        if (total) memmove(scratch, src, total);
        carry = total;
    }
    if (carry) memcpy(dst, scratch, carry);
    _mm_sfence();
}

}  // namespace wfs
