/* tests/compat/ref_malloc_pad.h -- test infrastructure.  Force-included (after <stdlib.h>) when the reference's OWN sources are
   compiled for the equivalence test: every malloc gets 64 spare bytes, which neutralises the reference's 2-float / 4-write
   heap overflow in MultiGrid2D::InitA (N2/MultiGrid2D.cpp:50-58) without touching its sources. */
#ifndef REF_MALLOC_PAD_H
#define REF_MALLOC_PAD_H
#include <stdlib.h>
static inline void* ref_padded_malloc(size_t n) { return malloc(n + 64); }
#define malloc(n) ref_padded_malloc(n)
#endif
