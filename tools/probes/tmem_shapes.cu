// Probe: which (lane, column) of tensor memory lands in which (thread, register) for the tcgen05.ld shapes.
// Tensor memory lanes 0..31 x columns 0..63 are filled with lane * 256 + column through the 32x32b shape (thread i <-> lane i,
// register r <-> column r); every other shape then reads 16 registers per thread and prints them.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tmem_shapes tmem_shapes.cu && ./tmem_shapes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define LD16(SHAPE_STR, OUT, ADDR)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned." SHAPE_STR ".b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t" \
                 "tcgen05.wait::ld.sync.aligned;"                                                                                      \
                 : "=r"(OUT[0]), "=r"(OUT[1]), "=r"(OUT[2]), "=r"(OUT[3]), "=r"(OUT[4]), "=r"(OUT[5]), "=r"(OUT[6]), "=r"(OUT[7]),      \
                   "=r"(OUT[8]), "=r"(OUT[9]), "=r"(OUT[10]), "=r"(OUT[11]), "=r"(OUT[12]), "=r"(OUT[13]), "=r"(OUT[14]), "=r"(OUT[15])   \
                 : "r"(ADDR) : "memory")

__global__ void probe(uint32_t *out)
{
    __shared__ uint32_t s_base;
    const int lane = threadIdx.x;
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" :: "r"((uint32_t)__cvta_generic_to_shared(&s_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_base;
    for (int c0 = 0; c0 < 64; c0 += 16) {
        uint32_t v[16];
        for (int r = 0; r < 16; r++) v[r] = (uint32_t)(lane * 256 + c0 + r);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                     :: "r"(base + c0), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
                        "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t o[16];
    LD16("32x32b.x16", o, base);
    for (int r = 0; r < 16; r++) out[(0 * 32 + lane) * 16 + r] = o[r];
    LD16("16x64b.x16", o, base);
    for (int r = 0; r < 16; r++) out[(1 * 32 + lane) * 16 + r] = o[r];
    LD16("16x128b.x8", o, base);
    for (int r = 0; r < 16; r++) out[(2 * 32 + lane) * 16 + r] = o[r];
    LD16("16x256b.x4", o, base);
    for (int r = 0; r < 16; r++) out[(3 * 32 + lane) * 16 + r] = o[r];
    LD16("16x64b.x16", o, base + (16u << 16));
    for (int r = 0; r < 16; r++) out[(4 * 32 + lane) * 16 + r] = o[r];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" :: "r"(base) : "memory");
}

int main()
{
    uint32_t *d, h[5 * 32 * 16];
    cudaMalloc(&d, sizeof h);
    cudaMemset(d, 0xff, sizeof h);
    probe<<<1, 32>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status: %s\n", cudaGetErrorString(e));
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *names[5] = { "32x32b.x16", "16x64b.x16", "16x128b.x8", "16x256b.x4", "16x64b.x16 @lane16" };
    for (int s = 0; s < 5; s++) {
        printf("== %s  (entries are lane:column)\n", names[s]);
        for (int t = 0; t < 32; t++) {
            printf("t%02d:", t);
            for (int r = 0; r < 16; r++) { uint32_t v = h[(s * 32 + t) * 16 + r]; printf(" %2u:%-2u", v >> 8, v & 255); }
            printf("\n");
        }
    }
    return 0;
}
