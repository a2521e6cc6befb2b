#define WFS_EXPAND_NAME expand_records_avx2
#define WFS_VEC __m256i
#define WFS_VSET1_16(x) _mm256_set1_epi16(x)
#define WFS_VSTOREU(p, v) _mm256_storeu_si256(reinterpret_cast<__m256i *>(p), v)
#define WFS_VLOADU(p) _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p))
#define WFS_VSTREAM(p, v) _mm256_stream_si256(reinterpret_cast<__m256i *>(p), v)
#include "expand_impl.inc"
