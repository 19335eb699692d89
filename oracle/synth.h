/* oracle/synth.h -- TEST INFRASTRUCTURE (not product code).
 *
 * Counter-based synthetic sequence generator shared by the oracle driver (refdrv.cpp), the C
 * restatement (mb_oracle.c) and, re-stated in numpy, by bench.py / tests (which must not import
 * oracle/ on the product path).  Residue p of sequence k of stream `which` (0 = input, 1 = output)
 * under seed `seed` is
 *
 *     1 + splitmix64( seed * 0x9E3779B97F4A7C15 + (2k + which) * 0xD1B54A32D192ED03 + p ) % nSym
 *
 * i.e. tokens are iid uniform on 1..nSym (token 0 is the reference's epsilon, src/eval.h:25, and
 * never appears in data).  This replaces the std::mt19937 stream SURVEY.md section 8(d) used, so
 * that the identical batch can be regenerated in C, C++ and numpy without shipping sequences.
 */
#ifndef MB_ORACLE_SYNTH_H
#define MB_ORACLE_SYNTH_H
#include <stdint.h>

static inline uint64_t mb_splitmix64 (uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

static inline int mb_synth_token (uint64_t seed, uint64_t k, int which, uint64_t p, int nSym) {
  const uint64_t base = seed * 0x9E3779B97F4A7C15ULL + (2 * k + (uint64_t) which) * 0xD1B54A32D192ED03ULL + p;
  return 1 + (int) (mb_splitmix64 (base) % (uint64_t) nSym);
}

#endif
