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

#ifdef _OPENMP
#include <omp.h>
#endif

/* thread count of the OpenMP build (the timed CPU baseline).  OMP_NUM_THREADS is read once when libgomp loads, and
 * torch.distributed.run exports OMP_NUM_THREADS=1 to its workers, so bench.py sets the count explicitly. */
void swref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0)
        omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int swref_get_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int swref_has_openmp(void)
{
#ifdef _OPENMP
    return 1;
#else
    return 0;
#endif
}
