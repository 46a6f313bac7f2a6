// tcgen05 operand-fetch microbenchmark + semantics probe (diagnostic, sm_100a).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
//
// (1) cycles per tcgen05.mma (M=128, K=16, SS operands) for SWIZZLE_NONE / 32B / 64B / 128B K-major
//     layouts and several N: tells how fast the tensor pipe can fetch A/B from shared memory.
// (2) does a swizzled K-major A operand work when its start address is shifted by an arbitrary
//     number of rows (the implicit-GEMM "tap = row shift" trick), and with which base_offset?
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../brats2019_b200/csrc/common.cuh"

using namespace b200;

__device__ __forceinline__ uint64_t desc_full(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_off & 7) << 49;
    d |= (uint64_t)(layout & 7) << 61;
    return d;
}

struct Cfg {
    int layout;      // 0 none, 6 sw32, 4 sw64, 2 sw128
    int rowbytes;    // 16 (none: chunk planes), 32, 64, 128
    int N;
    int iters;
    int shift_rows;  // A start shifted by this many rows
    int base_mode;   // 0: base_offset = 0, 1: base_offset = (start >> 7) & 7
    int k0;          // first K element (multiple of 16)
    int a_mn_major;  // timing variant
};

// smem: A region 512 rows * rowbytes (or planes), B region 256 rows * rowbytes
__global__ void __launch_bounds__(128, 1) probe_kernel(Cfg c, const uint8_t* a_img, const uint8_t* b_img, int a_bytes,
                                                        int b_bytes, float* out, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tslot;
    uint8_t* sa = smem;
    uint8_t* sb = smem + 96 * 1024;
    for (int i = threadIdx.x * 16; i < a_bytes; i += blockDim.x * 16) *(uint4*)(sa + i) = *(const uint4*)(a_img + i);
    for (int i = threadIdx.x * 16; i < b_bytes; i += blockDim.x * 16) *(uint4*)(sb + i) = *(const uint4*)(b_img + i);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(&tslot, 256);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tslot;
    if (threadIdx.x == 0) {
        const int Mmma = c.a_mn_major >= 2 ? c.base_mode : 128;       // MN-major timing variant: M passed in base_mode
        const uint32_t idesc = c.a_mn_major >= 2 ? make_idesc(Mmma, c.N, 1, 1) : make_idesc(128, c.N, c.a_mn_major, 0);
        uint32_t a_addr, b_addr, lbo_a, sbo_a, lbo_b, sbo_b;
        if (c.a_mn_major == 3) {
            // line-marching weight gradient (wgrad_line.cuh): block stride = one line plane (rowbytes field), start
            // address shifted by shift_rows rows of 16 B
            a_addr = smem_u32(sa) + c.shift_rows * 16; lbo_a = 128; sbo_a = c.rowbytes;
            b_addr = smem_u32(sb) + c.shift_rows * 16; lbo_b = 128; sbo_b = c.rowbytes;
        } else if (c.a_mn_major == 2) {
            // both operands MN-major, SWIZZLE_NONE, as the weight-gradient GEMM reads them: planes [8-ch block][row][16 B],
            // LBO = 128 B between 8-row K groups, SBO = plane stride between 8-channel blocks (64-row planes)
            a_addr = smem_u32(sa); lbo_a = 128; sbo_a = 64 * 16;
            b_addr = smem_u32(sb); lbo_b = 128; sbo_b = 64 * 16;
        } else if (c.layout == 0) {
            // chunk planes: [chunk][row][16B]; plane stride 512 rows * 16
            a_addr = smem_u32(sa) + c.shift_rows * 16 + (c.k0 / 8) * 512 * 16;
            lbo_a = 512 * 16; sbo_a = 128;
            b_addr = smem_u32(sb) + (c.k0 / 8) * 256 * 16;
            lbo_b = 256 * 16; sbo_b = 128;
        } else {
            a_addr = smem_u32(sa) + c.shift_rows * c.rowbytes + c.k0 * 2;
            lbo_a = 16; sbo_a = 8 * c.rowbytes;
            b_addr = smem_u32(sb) + c.k0 * 2;
            lbo_b = 16; sbo_b = 8 * c.rowbytes;
        }
        const uint32_t boff = (c.base_mode && c.a_mn_major < 2) ? ((a_addr >> 7) & 7) : 0;
        const uint64_t ad = desc_full(a_addr, lbo_a, sbo_a, c.layout, boff);
        const uint64_t bd = desc_full(b_addr, lbo_b, sbo_b, c.layout, 0);
        long long t0 = clock64();
        for (int i = 0; i < c.iters; ++i) umma_bf16(tm, ad, bd, idesc, i > 0);
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) cycles[0] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (blockIdx.x == 0 && out) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int c0 = 0; c0 < c.N; c0 += 16) {
            float v[16];
            tmem_ld16(tm + ((uint32_t)(warp * 32) << 16) + c0, v);
            for (int i = 0; i < 16; ++i) out[(warp * 32 + lane) * 256 + c0 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 256);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return (uint16_t)(u >> 16);
}
static int aval(int r, int k) { return ((r * 7 + k * 3) % 5) - 2; }
static int bval(int n, int k) { return ((n * 5 + k * 11) % 7) - 3; }

// physical byte offset of logical (row r, byte j within row) for the layout
static size_t phys(int layout, int rowbytes, int rows_per_plane, int r, int j) {
    if (layout == 0) return (size_t)(j / 16) * rows_per_plane * 16 + (size_t)r * 16 + (j % 16);
    size_t a = (size_t)r * rowbytes + j;
    int bits = layout == 6 ? 1 : layout == 4 ? 2 : 3;
    size_t x = (a >> 7) & ((1u << bits) - 1);
    return a ^ (x << 4);
}

int main() {
    const int AR = 512, BRW = 256;
    uint8_t *d_a, *d_b;
    float* d_out;
    long long* d_cyc;
    cudaMalloc(&d_a, 96 * 1024); cudaMalloc(&d_b, 64 * 1024);
    cudaMalloc(&d_out, 128 * 256 * 4); cudaMalloc(&d_cyc, 8);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    struct L { int layout, rowbytes; const char* name; } layouts[] = {{0, 16, "NONE "}, {6, 32, "SW32 "}, {4, 64, "SW64 "}, {2, 128, "SW128"}};
    // ---- (2) correctness with row-shifted starts ----
    printf("== shifted-start correctness (M=128,N=16,K=16): max |D - ref| ==\n");
    for (auto& L : layouts) {
        const int kel = L.layout == 0 ? 16 : L.rowbytes / 2;     // K elements stored per row
        std::vector<uint8_t> ha(96 * 1024, 0), hb(64 * 1024, 0);
        for (int r = 0; r < AR; ++r)
            for (int k = 0; k < kel; ++k) {
                uint16_t v = f2bf((float)aval(r, k));
                memcpy(&ha[phys(L.layout, L.rowbytes, AR, r, k * 2)], &v, 2);
            }
        for (int n = 0; n < BRW; ++n)
            for (int k = 0; k < kel; ++k) {
                uint16_t v = f2bf((float)bval(n, k));
                memcpy(&hb[phys(L.layout, L.rowbytes, BRW, n, k * 2)], &v, 2);
            }
        cudaMemcpy(d_a, ha.data(), ha.size(), cudaMemcpyHostToDevice);
        cudaMemcpy(d_b, hb.data(), hb.size(), cudaMemcpyHostToDevice);
        for (int k0 = 0; k0 < kel; k0 += 16) {
            if (k0 > 16) continue;
            for (int shift : {0, 1, 2, 3, 5, 8, 11, 130}) {
                for (int bm = 0; bm < 2; ++bm) {
                    if (L.layout == 0 && bm == 1) continue;
                    Cfg c{L.layout, L.rowbytes, 16, 1, shift, bm, k0, 0};
                    cudaMemset(d_out, 0, 128 * 256 * 4);
                    probe_kernel<<<1, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, d_out, d_cyc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%s shift %d: CUDA error %s\n", L.name, shift, cudaGetErrorString(e)); return 1; }
                    std::vector<float> ho(128 * 256);
                    cudaMemcpy(ho.data(), d_out, ho.size() * 4, cudaMemcpyDeviceToHost);
                    double maxerr = 0;
                    for (int m = 0; m < 128; ++m)
                        for (int n = 0; n < 16; ++n) {
                            double ref = 0;
                            for (int k = 0; k < 16; ++k) ref += aval(shift + m, k0 + k) * bval(n, k0 + k);
                            maxerr = std::max(maxerr, fabs(ref - ho[m * 256 + n]));
                        }
                    printf("%s k0=%2d shift=%3d base_offset=%s : max err %g %s\n", L.name, k0, shift,
                           bm ? "(addr>>7)&7" : "0          ", maxerr, maxerr == 0 ? "OK" : "WRONG");
                }
            }
        }
    }
    // ---- (1) timing ----
    printf("== cycles per tcgen05.mma (M=128,K=16), 148 CTAs concurrently, 2000 back-to-back MMAs ==\n");
    for (auto& L : layouts)
        for (int N : {16, 32, 48, 64, 96, 128, 192, 256}) {
            Cfg c{L.layout, L.rowbytes, N, 2000, 0, 0, 0, 0};
            probe_kernel<<<148, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, nullptr, d_cyc);
            cudaDeviceSynchronize();
            long long cyc;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d : %7.1f cycles/MMA  (floor N/2 = %d)\n", L.name, N, (double)cyc / 2000, N / 2);
        }
    printf("== both operands MN-major SWIZZLE_NONE (weight-gradient GEMM form), K=16: cycles per MMA ==\n");
    for (int M : {64, 128})
        for (int N : {16, 48, 96, 128, 144, 192, 256}) {
            Cfg c{0, 16, N, 2000, 0, M, 0, 2};
            probe_kernel<<<148, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, nullptr, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("MN-major M=%d N=%d: CUDA error %s\n", M, N, cudaGetErrorString(e)); return 1; }
            long long cyc;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("MN/MN M=%3d N=%3d : %7.1f cycles/MMA  (N/2 = %d)\n", M, N, (double)cyc / 2000, N / 2);
        }
    printf("== MN/MN M=64 N=144 (wgrad_line.cuh operands): block stride x start-row shift, cycles per MMA ==\n");
    for (int sbo : {1024, 2048, 2080, 2176, 2304})
        for (int shift : {0, 1, 4, 7, 8}) {
            Cfg c{0, sbo, 144, 2000, shift, 64, 0, 3};
            probe_kernel<<<148, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, nullptr, d_cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("line form sbo=%d shift=%d: CUDA error %s\n", sbo, shift, cudaGetErrorString(e)); return 1; }
            long long cyc;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("line form SBO=%4d B start shift=%d rows : %7.1f cycles/MMA\n", sbo, shift, (double)cyc / 2000);
        }
    printf("== shifted start (3 rows) timing, N=16/64 ==\n");
    for (auto& L : layouts)
        for (int N : {16, 64}) {
            Cfg c{L.layout, L.rowbytes, N, 2000, 3, 1, 0, 0};
            probe_kernel<<<148, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, nullptr, d_cyc);
            cudaDeviceSynchronize();
            long long cyc;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d shift3: %7.1f cycles/MMA\n", L.name, N, (double)cyc / 2000);
        }
    printf("== K-major SWIZZLE_NONE chunk planes, A start shifted by s rows (conv_gemm.cuh taps), N = 64 / 128 ==\n");
    for (int N : {64, 128})
        for (int shift : {0, 1, 2, 3, 4, 8, 19, 36}) {
            Cfg c{0, 16, N, 2000, shift, 0, 0, 0};
            probe_kernel<<<148, 128, 200 * 1024>>>(c, d_a, d_b, 96 * 1024, 64 * 1024, nullptr, d_cyc);
            cudaDeviceSynchronize();
            long long cyc;
            cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
            printf("NONE  N=%3d shift=%2d: %7.1f cycles/MMA\n", N, shift, (double)cyc / 2000);
        }
    return 0;
}
