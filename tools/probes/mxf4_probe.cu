// Compile-only probe (nvcc -c, never linked into libdmp2.so): which block-scaled tcgen05 forms does ptxas accept for
// sm_100a?  Preparation for a conv mode whose correction terms are FP4 (kind::mxf4 / mxf4nvf4, K = 64 per instruction).
#include <stdint.h>
__global__ void probe(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t sfa, uint32_t sfb, uint64_t sdesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf4nvf4.block_scale.scale_vec::4X [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb) : "memory");
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale.scale_vec::1X [%0], %1, %2, %3, [%4], [%5], p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(sfa), "r"(sfb) : "memory");
    // scale factors: shared memory -> tensor memory
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(sfa), "l"(sdesc) : "memory");
    asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(sfa), "l"(sdesc) : "memory");
}
