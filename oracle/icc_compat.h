/* TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Force-included when the unmodified reference /root/reference/spmv.cpp is
 * compiled with g++ instead of icpc.  The reference uses four Intel-compiler
 * spellings (KNL era); everything else is standard AVX-512F + OpenMP.
 * Recipe recorded in SURVEY.md section 8(c). */
#ifndef CVR_ORACLE_ICC_COMPAT_H
#define CVR_ORACLE_ICC_COMPAT_H
#include <immintrin.h>
#ifndef _MM_SCALE_8
#define _MM_SCALE_8 8
#endif
#ifndef _MM_SCALE_4
#define _MM_SCALE_4 4
#endif
/* icc: gather 8 doubles using the low 8 of 16 int32 indices */
#define _mm512_i32logather_pd(idx, base, scale) \
    _mm512_i32gather_pd(_mm512_castsi512_si256(idx), base, scale)
/* icc/KNC: permute the four 128-bit quarters of a zmm register */
#define _mm512_permute4f128_epi32(v, perm) _mm512_shuffle_i32x4(v, v, perm)
#endif
