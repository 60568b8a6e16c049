/*
 * oracle/mg_oracle.c -- TEST INFRASTRUCTURE, not product code.
 *
 * CPU restatement ("port") of the reference multigrid hot path in plain C: the operators and
 * V-cycle / FMG drivers of NOCUDA_TESI for the 1D equation, the 2D Lyapunov problem and the 3D
 * Poisson problem, in float and double, with the reference's residual (REF_COMPAT) and the
 * sign-corrected one (CORRECTED).  Every function in mg_oracle_impl.h cites the reference
 * file:line it follows.
 *
 * Parity pin: tests/test_oracle.py checks this file bit-for-bit against (a) oracle/_ref/
 * libmg_ref.so = the reference itself compiled unmodified, whenever that library is present,
 * and (b) the golden vectors in tests/golden/ that tests/golden/make_golden.py generated from it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (pde_multigrid_b200) has no CPU path.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math mg_oracle.c -o libmg_oracle.so -lm
 */
#include <math.h>
#include <stddef.h>
#include <stdlib.h>

static float orc_exp_f32(float x) { return expf(x); }
static double orc_exp_f64(double x) { return exp(x); }

#define REAL float
#define SFX _f32
#include "mg_oracle_impl.h"
#undef REAL
#undef SFX

#define REAL double
#define SFX _f64
#include "mg_oracle_impl.h"
#undef REAL
#undef SFX
