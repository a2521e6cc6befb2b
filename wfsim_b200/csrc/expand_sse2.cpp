#define WFS_EXPAND_NAME expand_records_sse2
#define WFS_VEC __m128i
#define WFS_VSET1_16(x) _mm_set1_epi16(x)
#define WFS_VSTOREU(p, v) _mm_storeu_si128(reinterpret_cast<__m128i *>(p), v)
#define WFS_VLOADU(p) _mm_loadu_si128(reinterpret_cast<const __m128i *>(p))
#define WFS_VSTREAM(p, v) _mm_stream_si128(reinterpret_cast<__m128i *>(p), v)
#include "expand_impl.inc"
