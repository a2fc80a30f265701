// tma.cuh -- TMA (cp.async.bulk.tensor) staging helpers shared by the fused 2D kernels (acou_vd_fused.cu, ela_fused.cu).
// A CTA's elected thread arms one mbarrier with the byte count of the tile's working set and issues one 2D box per staged
// array; the boxes are described by CUtensorMaps of whole padded planes (zeros outside the plane), built on the host and passed
// as __grid_constant__ kernel parameters.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

namespace swb {

// tensor map of a 2D row-major plane of `rows` rows with row pitch `ld` elements (dtype SWB_F32 / SWB_F64), box = box_w x box_h
// elements, no swizzle, zero fill outside the plane.  Host side (engine_core.cu); throws swb::Error.
void make_tmap_2d(CUtensorMap *out, int dtype, const void *base, long long ld, long long rows, int box_w, int box_h);
// false when SWB_NO_PDL is set: the step kernels are then launched with full stream serialization
bool pdl_enabled();

#ifdef __CUDACC__
// Programmatic dependent launch (consecutive step kernels of a sweep): a kernel launched with launch_pdl() may start while its
// predecessor on the stream drains; it must call pdl_wait() before it touches anything the predecessor writes (everything it requests
// before that -- material arrays -- overlaps the predecessor's tail and its own launch latency).  pdl_trigger() lets the successor's
// CTAs be scheduled as soon as every CTA of this grid has started.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile("{\n"
                 ".reg .pred P1;\n"
                 "LAB_WAIT:\n"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
                 "@P1 bra DONE;\n"
                 "bra LAB_WAIT;\n"
                 "DONE:\n"
                 "}\n" ::"r"(smem_u32(bar)),
                 "r"(parity)
                 : "memory");
}
// one box: columns c0 .. c0 + box_w - 1, rows c1 .. c1 + box_h - 1 of the plane -> dense rows of box_w elements at dst (128-byte aligned)
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(smem_u32(dst)), "l"(map), "r"(c0),
                 "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
#endif

} // namespace swb
