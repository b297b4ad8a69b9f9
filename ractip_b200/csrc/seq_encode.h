// seq_encode.h -- sequence letters -> the byte code the kernels read.
// low 3 bits: 0 N/other, 1 A, 2 C, 3 G, 4 U (T folds like U, ViennaRNA's
// encoding); bit 3 set when the upper-cased letter is not one of A C G U, so
// that special-hairpin list matching (a string compare in ViennaRNA) cannot
// match such a window.
#ifndef RP_SEQ_ENCODE_H
#define RP_SEQ_ENCODE_H
#include <stdint.h>
namespace rp {
inline uint8_t encode_base(char ch) {
  switch (ch) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'G': case 'g': return 3;
    case 'U': case 'u': return 4;
    case 'T': case 't': return 4 | 8;
    default: return 0 | 8;
  }
}
}  // namespace rp
#endif
