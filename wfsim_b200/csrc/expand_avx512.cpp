#define WFS_EXPAND_NAME expand_records_avx512
#define WFS_VEC __m512i
#define WFS_VSET1_16(x) _mm512_set1_epi16(x)
#define WFS_VSTOREU(p, v) _mm512_storeu_si512(reinterpret_cast<void *>(p), v)
#define WFS_VLOADU(p) _mm512_loadu_si512(reinterpret_cast<const void *>(p))
#define WFS_VSTREAM(p, v) _mm512_stream_si512(reinterpret_cast<__m512i *>(p), v)
#include "expand_impl.inc"
