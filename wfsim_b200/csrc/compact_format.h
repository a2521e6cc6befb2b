// Compact record transport format shared by the pack kernel (backend.cu), the host expanders
// (expand_*.cpp) and the transport plumbing (transport.cuh).  Plain C++ (no CUDA types).
#pragma once
#include <stdint.h>

namespace wfs {

constexpr int kBlocksPerRecord = 14;    // ceil(110 / 8)

struct CompactHdr {
    int64_t time;
    int32_t pulse_length;
    int16_t channel;
    int16_t record_i;
    uint32_t boff;       // index of the record's first block in the block stream
    uint16_t mask;       // bit b: block b (samples [8b, 8b+8)) is in the stream
    uint16_t length;
};
static_assert(sizeof(CompactHdr) == 24, "CompactHdr layout");

// one expander per instruction set; transport.cu dispatches on the CPU it runs on
void expand_records_sse2(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);
void expand_records_avx2(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);
void expand_records_avx512(const CompactHdr *, const uint8_t *, int64_t, int64_t, uint8_t *, int16_t, int16_t);

}  // namespace wfs
