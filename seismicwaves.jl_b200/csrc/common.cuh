// common.cuh -- shared plumbing for libswb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <stdexcept>
#include <atomic>
#include "../../include/swb200.h"

namespace swb {

// ---- error handling -----------------------------------------------------------------------
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string &msg);
extern std::atomic<long long> g_launches;

#define SWB_CUDA(expr)                                                                                       \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            throw ::swb::Error(SWB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +      \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")");           \
    } while (0)

#define SWB_REQUIRE(cond, msg)                                     \
    do {                                                           \
        if (!(cond))                                               \
            throw ::swb::Error(SWB_ERR_ARG, std::string(msg));     \
    } while (0)

// wrap a C-ABI body: translate exceptions to status codes
#define SWB_API_BEGIN try {
#define SWB_API_END                                 \
    return SWB_OK;                                  \
    }                                               \
    catch (const ::swb::Error &e) {                 \
        ::swb::set_last_error(e.what());            \
        return e.code;                              \
    }                                               \
    catch (const std::bad_alloc &) {                \
        ::swb::set_last_error("host out of memory");\
        return SWB_ERR_NOMEM;                       \
    }                                               \
    catch (const std::exception &e) {               \
        ::swb::set_last_error(e.what());            \
        return SWB_ERR_STATE;                       \
    }

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
inline void check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        throw Error(SWB_ERR_CUDA, std::string("launch of ") + what + ": " + cudaGetErrorString(e));
}

// ---- arithmetic policy ----------------------------------------------------------------------
// T  = storage type (the reference's type parameter).
// CT = type the Float64-literal expressions evaluate in: double for the reference-faithful path
//      (src/utils/fdgen.jl:61-63,131), float for SWB_FLAG_FAST_F32.
// The library is compiled with -fmad=false so that a*b+c rounds twice like the reference's
// Julia code (LLVM never contracts without fastmath); fast mode uses explicit fma where it pays.
template <class T, class CT>
struct Arith {
    typedef T store_t;
    typedef CT comp_t;
};

// Fornberg weights (src/utils/fdgen.jl:11-47), computed on the host in double at library load.
struct FdWeights {
    double d1o2[2]; // deriv 1, order 2: offsets {0, +1}
    double d2o2[3]; // deriv 2, order 2: offsets {-1, 0, +1}
    double d1o4[4]; // deriv 1, order 4: offsets {-1, 0, +1, +2}
};
const FdWeights &fd_weights();

// column-major linear index, 0-based inputs
__host__ __device__ inline size_t lin2(int64_t i, int64_t j, int64_t n1) { return (size_t)j * (size_t)n1 + (size_t)i; }
__host__ __device__ inline size_t lin3(int64_t i, int64_t j, int64_t k, int64_t n1, int64_t n2)
{
    return ((size_t)k * (size_t)n2 + (size_t)j) * (size_t)n1 + (size_t)i;
}

inline unsigned cdiv(long long a, long long b) { return (unsigned)((a + b - 1) / b); }

// ---- device helpers shared by the physics files ---------------------------------------------------

// C-PML memory-variable update of a first derivative (src/utils/fdgen.jl:137-161):
//   psi <- b*psi + a*D  (b*psi in T, a*D in CT), returns D + psi_new (psi re-read after rounding to T)
template <class T, class CT>
__device__ __forceinline__ CT cpml_apply(CT D, T a, T b, T psi_old, T &psi_new)
{
    T bp = b * psi_old;
    psi_new = (T)((CT)bp + (CT)a * D);
    return D + (CT)psi_new;
}

} // namespace swb
