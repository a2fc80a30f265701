/*
 * ORACLE (test infrastructure, NOT product code).
 * CPU restatement of SeismicWaves.jl's CPU backend kernels; see swref_acoustic.h /
 * swref_elastic.h for the per-kernel citations.  Built by oracle/Makefile into
 * oracle/libswref.so (serial, bit-reproducible) and oracle/libswref_omp.so (OpenMP,
 * used only as the timed CPU baseline).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load these libraries.
 */
#include <stddef.h>
#include <math.h>

#define REAL float
#define FN(x) x##_f32
#include "swref_acoustic.h"
#ifdef SWREF_HAVE_ELASTIC
#include "swref_elastic.h"
#endif
#undef REAL
#undef FN

#define REAL double
#define FN(x) x##_f64
#include "swref_acoustic.h"
#ifdef SWREF_HAVE_ELASTIC
#include "swref_elastic.h"
#endif
#undef REAL
#undef FN

int swref_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}
