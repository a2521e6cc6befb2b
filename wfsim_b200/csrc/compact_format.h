// Compact record transport format shared by the pack kernel (backend.cu), the host expanders
// (expand_*.cpp) and the transport plumbing (transport.cuh).  Plain C++ (no CUDA types).
#pragma once
#include <stdint.h>

namespace wfs {

constexpr int kBlocksPerRecord = 28;    // 4-sample (8-byte) blocks: ceil(110 / 4)
constexpr int kBlockBytes = 8;
constexpr int kSamplesPerRecord = 110;

struct CompactHdr {
    int64_t time;
    int32_t pulse_length;
    int16_t channel;
    int16_t record_i;
    uint32_t boff;       // index of the record's first block in the block stream
    uint32_t mask;       // bit b: block b (samples [4b, 4b+4)) is in the stream
    // `length` is not carried: min(pulse_length - 110 * record_i, 110) (strax_interface.py:432)
};

static inline uint32_t compact_length(const CompactHdr &h) {
    const int64_t rest = (int64_t)h.pulse_length - (int64_t)kSamplesPerRecord * h.record_i;
    return (uint32_t)(rest < 0 ? 0 : rest > kSamplesPerRecord ? kSamplesPerRecord : rest);
}
static_assert(sizeof(CompactHdr) == 24, "CompactHdr layout");

// one expander per instruction set; transport.cu dispatches on the CPU it runs on
void expand_records_sse2(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);
void expand_records_avx2(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);
void expand_records_avx512(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);

}  // namespace wfs
