// Hardware probe (test hook, not on the product path): does a tcgen05.mma A-operand descriptor with a start
// address that is not 1024-byte aligned and a stride between 8-row groups that is not a multiple of 1024 bytes
// still resolve the 128-byte swizzle the way TMA wrote the tile?  If it does, a 3x3 tap is a VIEW (start offset
// + group stride) of one halo tile loaded once, instead of nine TMA im2col loads.
#pragma once
#include "conv_umma.cuh"

namespace dtraj {

// smem tile: `rows` pixel rows of 128 bytes, 128B-swizzled on absolute address bits (what TMA produces in a
// 1024-aligned buffer).  Element (r, 0) = r, everything else 0.  B = 16 x 8 K-major, B[0][0] = 1.
// D[m][0] = id of the row the tensor core fetched for output row m (or 0 / garbage if the swizzle phase is off).
__global__ void __launch_bounds__(128) k_probe_view(int rows, int start_row, int sbo_bytes, int base_off_mode, float* out) {
    extern __shared__ __align__(1024) uint8_t probe_sm[];
    const uint32_t base = (ptx::smem_u32(probe_sm) + 1023u) & ~1023u;
    uint8_t* a = probe_sm + (base - ptx::smem_u32(probe_sm));
    uint8_t* b = a + 32768;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 32768 / 4 + 2048 / 4; i += blockDim.x) reinterpret_cast<float*>(a)[i] = 0.f;
    __syncthreads();
    for (int r = threadIdx.x; r < rows; r += blockDim.x)
        *reinterpret_cast<float*>(a + r * 128 + ((0u ^ (uint32_t)(r & 7)) << 4)) = (float)r;
    if (threadIdx.x == 0) *reinterpret_cast<float*>(b) = 1.f;     // row 0, chunk 0 ^ 0, element 0
    ptx::fence_proxy_async();
    if (threadIdx.x < 32) {
        if (ptx::elect_one()) { ptx::mbar_init(ptx::smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        const uint32_t a_addr = base + (uint32_t)start_row * 128u;
        uint64_t ad = (uint64_t)((a_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)((uint32_t)sbo_bytes >> 4) << 32) |
                      ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
        if (base_off_mode) ad |= (uint64_t)((a_addr >> 7) & 7u) << 49;
        const uint64_t bd = umma_desc_sw128(base + 32768u);
        ptx::mma_tf32(tmem, ad, bd, umma_idesc_tf32(16), 0u);
        ptx::tc_commit(ptx::smem_u32(&bar));
    }
    ptx::mbar_wait(&g_umma_error, ptx::smem_u32(&bar), 0);
    ptx::tc_fence_after();
    uint32_t v[32];
    ptx::tmem_ld32(tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16), v);
    ptx::tmem_ld_wait();
    out[threadIdx.x] = __uint_as_float(v[0]);
    ptx::tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32u) : "memory");
}

// Second probe (for the next step of the conv kernel): does a TMA box over PERMUTED dimensions {c, x, image, y} of an
// NHWC fp16 map -- global strides not in increasing order -- deliver the halo rows of two 8x8 images interleaved in
// shared memory as [halo row (10)][image (2)][halo column (10)][64 channels], with out-of-bounds rows / columns
// zero-filled?  (Interleaved rows make consecutive 8-pixel row groups always 10 pixels apart, which is what a tap VIEW of
// a two-image tile needs.)  The map is encoded on the host; out[r] = first channel of shared-memory row r.
__global__ void __launch_bounds__(128) k_probe_tma_perm(const __grid_constant__ CUtensorMap map, int img0, float* out) {
    extern __shared__ __align__(1024) uint8_t probe_sm[];
    const uint32_t base = (ptx::smem_u32(probe_sm) + 1023u) & ~1023u;
    uint8_t* a = probe_sm + (base - ptx::smem_u32(probe_sm));
    __shared__ uint64_t bar;
    for (int i = threadIdx.x; i < 200 * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(a)[i] = -1.f;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    ptx::fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) {
        ptx::mbar_expect_tx(ptx::smem_u32(&bar), 200u * 128u);
        ptx::tma_load_4d(base, &map, ptx::smem_u32(&bar), 0, -1, img0, -1);
    }
    ptx::mbar_wait(&g_umma_error, ptx::smem_u32(&bar), 0);
    for (int r = threadIdx.x; r < 200; r += blockDim.x) out[r] = __half2float(*reinterpret_cast<const __half*>(a + r * 128));
}

}  // namespace dtraj
