// 3x3x3 implicit-GEMM convolution on CTA PAIRS: tcgen05.mma.cta_group::2 (M = 256 = two 128-row tiles, one per SM of a
// two-CTA cluster), the deep levels of the U-Net (128 -> 128 channels at 16^3, model.py:72-73 reached from the level-3
// residual blocks, and the data-gradient half of their backward).
//
// Why: at batch 2 the 128-channel level has only ~91 tiles of 128 rows, one per CTA, and every CTA needs the WHOLE
// 885 KB weight tensor once.  Measured with the probes build (tools/conv_deep_probe.py): weight stream alone 8.5 us
// (96 SMs x 885 KB = 10 TB/s, the L2 -> SM limit), MMAs alone 7.2 us, together 20.4 us - a four-deep ring of 37 KB weight
// stages cannot keep the tensor pipe fed at that latency.  A pair shares the B operand: each CTA loads only HALF of every
// weight stage (its 64 of the 128 output channels) and the hardware reads both halves, so the per-SM weight stream and
// the L2 traffic halve while the ring holds twice as many stages in the same shared memory.
//
// Same data layout, tile geometry, taps-as-start-address trick and epilogue as conv_gemm.cuh (MODE_K3, EPI_BF16, no
// kw-fold); the packed weights are the generic image packed with Nmma = Cout / 2 and two column jobs: job r is exactly
// the half CTA r of the pair needs, one contiguous bulk copy per stage.
//
// Pair protocol (rank 0 = leader issues every MMA):
//   * both CTAs run their own activation and weight producers into their own rings (same stage indices in lock step);
//   * full barriers: the leader's have two arrivals - its own producer's expect_tx and a relay arrive from the peer,
//     whose warp 2 waits on the peer's local full barrier and then arrives remotely (cp.async.bulk, unlike the tensor
//     form of TMA, cannot signal an mbarrier in another CTA);
//   * empty barriers and the accumulator-full barrier: tcgen05.commit.cta_group::2 with multicast to both CTAs;
//   * accumulator-empty: on the leader, counting the epilogue threads of both CTAs (the peer's arrive remotely);
//   * each CTA's epilogue reads its own TMEM (its 128 rows x all Cout columns) and stores its own tile.
#pragma once
#include "conv_gemm.cuh"

namespace b200 {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {     // every thread of every CTA of the cluster
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the shared::cta address `local` in CTA `rank` of this cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {  // whole warp, same warp in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate));
}
// Arrive on the mbarrier at this shared-memory offset in BOTH CTAs of the pair once all MMAs issued so far completed.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// NFULL = Cout = N of the pair's MMA; each CTA holds NFULL / 2 columns of B.  grid = 2 * pairs, cluster (2, 1, 1).
template <int NFULL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kConvThreads, 1)
conv_pair_kernel(const __grid_constant__ ConvKParams p) {
    constexpr int NH = NFULL / 2;
    constexpr int CO = NFULL;
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int cta = blockIdx.x;                                   // row of stats_partial
    const int pair_tiles = (p.num_tiles + 1) >> 1;                // the pair takes tiles 2*tp and 2*tp + 1

    uint8_t* smem_x = smem + p.smem_x_off;
    uint8_t* smem_w = smem + p.smem_w_off;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.smem_bar_off);
    uint64_t* x_full = bars;
    uint64_t* x_empty = x_full + p.x_stages;
    uint64_t* w_full = x_empty + p.x_stages;
    uint64_t* w_empty = w_full + p.w_stages;
    uint64_t* t_full = w_empty + p.w_stages;
    uint64_t* t_empty = t_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);
    float* stat_smem = reinterpret_cast<float*>(tmem_slot + 4);  // [4*kEpiGroups warps][16]

    if (warp == 1 && lane == 0) {
        const uint32_t nfull = rank == 0 ? 2u : 1u;              // leader: own producer + the peer's relay
        for (int i = 0; i < p.x_stages; ++i) { mbar_init(&x_full[i], nfull); mbar_init(&x_empty[i], 1); }
        for (int i = 0; i < p.w_stages; ++i) { mbar_init(&w_full[i], nfull); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&t_full[i], 1); mbar_init(&t_empty[i], 2 * 128 * kEpiGroups); }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc_pair(tmem_slot, p.tmem_cols);
    if (p.stats_partial)
        for (int i = threadIdx.x; i < p.N * 16; i += blockDim.x) p.stats_partial[(size_t)cta * p.N * 16 + i] = 0.f;
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // the peer's barriers are initialised and its TMEM is allocated before anything crosses over
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int R = p.BD * p.MB;
    const uint32_t leader_bars = mapa_u32(smem_u32(bars), 0);    // the leader's barrier block in the cluster window

    if (warp == 0) {
        // ================= activation producer: this CTA's own tile =================
        int xs = 0; uint32_t xph = 0;
        for (int tp = pair; tp < pair_tiles; tp += npairs) {
            const int t = min(2 * tp + (int)rank, p.num_tiles - 1);          // odd tile count: the last peer tile is a
            TileCoord tc = decode_tile(p, t);                                // duplicate whose results are dropped
            for (int g = 0; g < p.KG; ++g) {
                mbar_wait(&x_empty[xs], xph ^ 1);
                if (elect_one()) {
                    mbar_arrive_expect_tx(&x_full[xs], p.x_stage_bytes);
                    const int chunk0 = g * (p.KC / 8);
                    uint8_t* dst = smem_x + (size_t)xs * p.x_stage_bytes;
                    const uint32_t bytes = (uint32_t)p.SRp * 16;
                    const size_t plane_stride = (size_t)p.src_a.plane_rows * 8;
                    const int nch = p.KC / 8;
                    for (int s = 0; s < p.nslices; ++s) {
                        const int dpi = (p.whole ? 0 : tc.d0 + 1) - 1 + s;
                        const long long row0 = ((long long)tc.n * (p.D + 2) + dpi) * p.SS + tc.q0 - p.halo_rows;
                        const __nv_bfloat16* ps = p.src_a.at(chunk0, row0);
                        uint8_t* pd = dst + (size_t)s * bytes;
                        for (int c = 0; c < nch; ++c, ps += plane_stride, pd += p.x_plane_bytes)
                            bulk_load_1d(pd, ps, bytes, &x_full[xs]);
                    }
                }
                __syncwarp();
                if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================= weight producer: this CTA's half (column job = rank) of every stage =================
        int ws = 0; uint32_t wph = 0;
        const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.wpacked) + (size_t)rank * p.KG * p.NTG * p.w_stage_bytes;
        for (int tp = pair; tp < pair_tiles; tp += npairs) {
            for (int g = 0; g < p.KG; ++g) {
                for (int tg = 0; tg < p.NTG; ++tg) {
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&w_full[ws], p.w_stage_bytes);
                        bulk_load_1d(smem_w + (size_t)ws * p.w_stage_bytes,
                                     wsrc + (size_t)(g * p.NTG + tg) * p.w_stage_bytes, p.w_stage_bytes, &w_full[ws]);
                    }
                    __syncwarp();
                    if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp == 2 && rank != 0) {
        // ================= peer: relay "my stage is full" to the leader's barriers =================
        int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
        const uint32_t l_x_full = leader_bars, l_w_full = leader_bars + 8u * (uint32_t)(2 * p.x_stages);
        for (int tp = pair; tp < pair_tiles; tp += npairs) {
            for (int g = 0; g < p.KG; ++g) {
                mbar_wait(&x_full[xs], xph);
                if (elect_one()) mbar_arrive_cluster(l_x_full + 8u * (uint32_t)xs);
                __syncwarp();
                for (int tg = 0; tg < p.NTG; ++tg) {
                    mbar_wait(&w_full[ws], wph);
                    if (elect_one()) mbar_arrive_cluster(l_w_full + 8u * (uint32_t)ws);
                    __syncwarp();
                    if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
                }
                if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
            }
        }
    } else if (warp == 2) {
        // ================= leader: MMA issuer for the pair (M = 256, N = NFULL) =================
        constexpr uint32_t idesc = make_idesc(256, NFULL, 0, 0);
        const uint32_t plane16 = p.x_plane_bytes >> 4;
        const uint64_t a_hi = ((uint64_t)(plane16 & 0x3FFF) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint64_t b_hi = ((uint64_t)((NH * 16) >> 4) << 16) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
        const uint32_t xbase16 = smem_u32(smem_x) >> 4, wbase16 = smem_u32(smem_w) >> 4;
        const uint32_t xstage16 = p.x_stage_bytes >> 4, wstage16 = p.w_stage_bytes >> 4;
        int xs = 0, ws = 0; uint32_t xph = 0, wph = 0;
        int it = 0;
        for (int tp = pair; tp < pair_tiles; tp += npairs, ++it) {
            const int as = it & 1;
            mbar_wait(&t_empty[as], ((it >> 1) & 1) ^ 1);
            tc_fence_after();
            for (int g = 0; g < p.KG; ++g) {
                mbar_wait(&x_full[xs], xph);
                tc_fence_after();
                const uint32_t xst16 = xbase16 + xs * xstage16;
                for (int tg = 0; tg < p.NTG; ++tg) {
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    const uint32_t wst16 = wbase16 + ws * wstage16;
                    const int* toff = &p.tap_off[tg * 9];
                    uint32_t dtm = tmem_base + (uint32_t)(as * R * NFULL);
                    const uint32_t first = (g | tg) == 0 ? 0u : 1u;
                    uint32_t tof[9];
#pragma unroll
                    for (int i = 0; i < 9; ++i) tof[i] = (uint32_t)toff[i];
                    const int BD = p.BD, MB = p.MB, SRp = p.SRp;
                    if (elect_one()) {
                        for (int dz = 0; dz < BD; ++dz) {
                            for (int mb = 0; mb < MB; ++mb, dtm += NFULL) {
                                const uint32_t a_run16 = xst16 + (uint32_t)(dz * SRp + mb * 128);
#pragma unroll
                                for (int tl = 0; tl < 9; ++tl) {
                                    umma_bf16_pair(dtm, a_hi | (uint64_t)(a_run16 + tof[tl]),
                                                   b_hi | (uint64_t)(wst16 + (uint32_t)(tl * 2 * NH)), idesc,
                                                   tl == 0 ? first : 1u);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (elect_one()) umma_commit_pair(&w_empty[ws]);
                    __syncwarp();
                    if (++ws == p.w_stages) { ws = 0; wph ^= 1; }
                }
                if (elect_one()) umma_commit_pair(&x_empty[xs]);
                __syncwarp();
                if (++xs == p.x_stages) { xs = 0; xph ^= 1; }
            }
            if (elect_one()) umma_commit_pair(&t_full[as]);
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ================= epilogue: this CTA's 128 rows x NFULL columns (as conv_gemm.cuh, MODE_K3 / EPI_BF16) =================
        const int grp = (warp - 4) >> 2;
        const int ew = (warp - 4) & 3;
        const int m = ew * 32 + lane;
        constexpr int GS = CO / 8;
        float ssum[8], ssq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        int cur_n = -1;
        const bool do_stats = p.stats_partial != nullptr;
        const uint32_t l_t_empty = leader_bars + 8u * (uint32_t)(2 * p.x_stages + 2 * p.w_stages + 2);

        auto flush_stats = [&](int n) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = warp_sum(ssum[i]); ssq[i] = warp_sum(ssq[i]); }
            const int w8 = warp - 4;
            if (lane == 0) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { stat_smem[w8 * 16 + i] = ssum[i]; stat_smem[w8 * 16 + 8 + i] = ssq[i]; }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiGroups) : "memory");
            if (w8 == 0 && lane < 16) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < 4 * kEpiGroups; ++k) v += stat_smem[k * 16 + lane];
                p.stats_partial[((size_t)cta * p.N + n) * 16 + lane] = v;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(128 * kEpiGroups) : "memory");
#pragma unroll
            for (int i = 0; i < 8; ++i) { ssum[i] = 0.f; ssq[i] = 0.f; }
        };

        int it = 0;
        for (int tp = pair; tp < pair_tiles; tp += npairs, ++it) {
            const int as = it & 1;
            const int traw = 2 * tp + (int)rank;
            const bool dummy = traw >= p.num_tiles;
            TileCoord tc = decode_tile(p, min(traw, p.num_tiles - 1));
            if (do_stats && !dummy && tc.n != cur_n) {
                if (cur_n >= 0) flush_stats(cur_n);
                cur_n = tc.n;
            }
            mbar_wait(&t_full[as], (it >> 1) & 1);
            tc_fence_after();
            for (int r = grp; r < R; r += kEpiGroups) {
                const int dz = r / p.MB, mb = r - dz * p.MB;
                const int q = tc.q0 + mb * 128 + m;
                const int dpo = (p.whole ? 0 : tc.d0 + 1) + dz;
                const int dq = p.by_SS.div(q);
                const int r2 = q - dq * p.SS;
                const int hp = p.by_Wp.div(r2);
                const int wp = r2 - hp * p.Wp;
                const int dp = dpo + dq;
                const bool valid = !dummy && (q < p.Q0 + p.QN) && dp >= 1 && dp <= p.D && hp >= 1 && hp <= p.H && wp >= 1 &&
                                   wp <= p.W;
                const long long orow = ((long long)tc.n * (p.D + 2) + dpo) * p.SS + q;
                const uint32_t trow = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)((as * R + r) * NFULL);
#pragma unroll
                for (int c0 = 0; c0 < CO; c0 += 16) {
                    float v[16];
                    tmem_ld16(trow + c0, v);
                    if (valid) {
                        if (do_stats) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const int gi = (c0 + i) / GS;
                                ssum[gi] += v[i];
                                ssq[gi] += v[i] * v[i];
                            }
                        }
                        const int ch = c0 >> 3;
                        if (p.residual.base) {
                            float f[8];
                            unpack_bf16x8(ld_nc_v4(p.residual.at(ch, orow)), f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] += f[i];
                            unpack_bf16x8(ld_nc_v4(p.residual.at(ch + 1, orow)), f);
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[8 + i] += f[i];
                        }
                        if (p.lrelu_out) {
#pragma unroll
                            for (int i = 0; i < 16; ++i) v[i] = lrelu(v[i]);
                        }
                        *reinterpret_cast<uint4*>(p.out.at(ch, orow)) = pack_bf16x8(v);
                        *reinterpret_cast<uint4*>(p.out.at(ch + 1, orow)) = pack_bf16x8(v + 8);
                    }
                }
            }
            tc_fence_before();
            if (rank == 0) mbar_arrive(&t_empty[as]);
            else mbar_arrive_cluster(l_t_empty + 8u * (uint32_t)as);
        }
        if (do_stats && cur_n >= 0) flush_stats(cur_n);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();          // neither CTA may exit (or free TMEM) while the other can still signal it or read its shared memory
    if (warp == 3) tmem_dealloc_pair(tmem_base, p.tmem_cols);
    if (p.gn_fin.mean != nullptr &&
        cta_draws_last_ticket(p.gn_fin.ticket, gridDim.x, reinterpret_cast<unsigned int*>(smem + 2048)))
        gn_stats_finalize_cta(p.stats_partial, (int)gridDim.x, p.N, p.gn_fin, reinterpret_cast<double*>(smem));
}

}  // namespace b200
