// Weight gradient of a k-tap conv on the F16 tensor pipe: wgrad_pp_kernel (tc_wgradp.cuh) with both operands carried as fp16
// hi/lo pieces of power-of-two-scaled tensors (see tc_convh.cuh for the arithmetic and the range argument).
//
//     G[(tap, n)][m] = sum_p Xs[p, (tap, n)] * dY[p, m]            rows = (tap, x channel), columns = dY channels
//
// Both tensors are [positions][channels] in memory, so the reduction index is the OUTER one for both operands:
//   A = Xs^T through tensor memory (thread = output row reads its column of a [32 pos][32 ch] box, as in wgrad_pp) -- now
//       scaled, split and packed as fp16 pairs along K: 16 + 16 columns per 32-position chunk instead of 32 + 32;
//   B = dY: wgrad_pp fed the fp32 tile to the tensor core MN-major.  Here converter warps TRANSPOSE it while splitting:
//       thread = dY channel reads its column of the raw [32 pos][32 ch] box (conflict-free) and writes one K-major row
//       (32 fp16 = 64 B, SWIZZLE_64B) of the hi tile and of the lo tile -- the operand layout conv_hp_kernel already uses.
// 12 kind::f16 MMAs per chunk (M = 256, N = h0 / h1, K = 16) instead of 24 kind::tf32 MMAs.
// With the tensor-pipe time halved the CONVERTERS set the pace (~300 instructions per thread per chunk against 960 MMA
// cycles): two warps share every converter row set (16 of the chunk's 32 positions each: 8 A warps + 10 B warps), and the
// sample-edge test is one ballot per chunk (lane j = position j) instead of a compare chain per element.
#pragma once
#include "tc_wgradp.cuh"
#include "tc_convh.cuh"

namespace bm {
namespace tc {

constexpr int WH_BK = 32, WH_STAGES = 4, WH_ASTAGES = 3, WH_THREADS = 672;        // 21 warps, see the kernel
constexpr int WH_RAW_BYTES = WP_MAX_BLKS * WP_BLK_BYTES;        // 20 KB: this CTA's dY boxes, raw fp32
constexpr int WH_BH_BYTES = WP_MAX_BLKS * 32 * WH_BK * 2;       // 10 KB: 160 K-major rows of 64 B (hi); the same again for lo
constexpr int WH_STAGE_BYTES = WH_RAW_BYTES + 2 * WH_BH_BYTES;  // 40 KB
constexpr int WH_ATILE_BYTES = 4 * WP_BLK_BYTES;                // 16 KB: X boxes of one chunk
constexpr int WH_SMEM_BYTES = WH_STAGES * WH_STAGE_BYTES + WH_ASTAGES * WH_ATILE_BYTES + 1024;
constexpr int WH_ACC_COLS = 320, WH_A_COLS = WH_BK;             // 16 packed hi columns + 16 packed lo columns

__device__ __forceinline__ void tmem_st8u(uint32_t taddr, const uint32_t* r) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

struct WgradHP {
    WgradPP c;
    const float* x_amax;
    const float* dy_amax;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(WH_THREADS, 1)
wgrad_hp_kernel(const __grid_constant__ CUtensorMap tmDY, const __grid_constant__ CUtensorMap tmX, const WgradHP hp) {
    const WgradPP& p = hp.c;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[WH_STAGES], conv_bar[WH_STAGES], empty_bar[WH_STAGES], tmem_full_bar;
    __shared__ __align__(8) uint64_t afull_bar[WH_ASTAGES], aempty_bar[WH_ASTAGES];
    __shared__ uint32_t tmem_base_smem;
    __shared__ int prior_error;

    if (threadIdx.x == 0) prior_error = p.err ? *reinterpret_cast<volatile int*>(p.err) : 0;
    __syncthreads();
    const bool skip = prior_error != 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    int pair = blockIdx.x >> 1;
    const int n_tile = pair % p.ntiles; pair /= p.ntiles;
    const int m_tile = pair % p.mtiles;
    const int ksl = pair / p.mtiles;
    const int it_begin = ksl * p.per_split;
    const int total = skip ? 0 : max(0, min(p.chunks, it_begin + p.per_split) - it_begin);
    const int n0 = n_tile * p.nt;                                  // first dY channel of the N tile
    // this CTA's dY channels: [n0 + rank*h0/2, +h0/2) and [n0 + h0 + rank*h1/2, +h1/2)
    const int blk0 = p.h0 / 64, blk1 = p.h1 / 64;                   // 32-channel blocks per CTA in each half
    const int nblk = blk0 + blk1;
    const uint32_t raw_bytes = (uint32_t)nblk * WP_BLK_BYTES;

    if (threadIdx.x == 0) {
        for (int s = 0; s < WH_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&conv_bar[s], 2 * (8 + 2 * nblk));            // one elected lane per converter warp of both CTAs (LEADER's copy)
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < WH_ASTAGES; ++s) {
            mbar_init(&afull_bar[s], 1);
            mbar_init(&aempty_bar[s], 8);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    if (warp == 0) {
        // ------------------------------------------------ TMA producer: this CTA's half of the dY tile, raw ------------
        if (lane == 0) {
            prefetch_tmap(&tmDY);
            for (int it = 0; it < total; ++it) {
                const int s = it % WH_STAGES;
                const uint32_t ph = (it / WH_STAGES) & 1;
                if (!mbar_wait(&empty_bar[s], ph ^ 1, p.err, 101)) break;
                const int p0 = (it_begin + it) * WH_BK;
                uint8_t* st = smem + s * WH_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], raw_bytes);
                for (int k = 0; k < blk0; ++k)
                    tma_load_2d(st + k * WP_BLK_BYTES, &tmDY, &full_bar[s], n0 + (int)rank * (p.h0 / 2) + 32 * k, p0);
                for (int k = 0; k < blk1; ++k)
                    tma_load_2d(st + (blk0 + k) * WP_BLK_BYTES, &tmDY, &full_bar[s],
                                n0 + p.h0 + (int)rank * (p.h1 / 2) + 32 * k, p0);
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------ MMA issuer (leader): whole warp loops, one elected lane issues -
        if (leader) {
            const uint32_t idesc0 = umma_idesc_f16(256, p.h0);
            const uint32_t idesc1 = p.h1 ? umma_idesc_f16(256, p.h1) : 0u;
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % WH_STAGES;
                const uint32_t ph = (it / WH_STAGES) & 1;
                ok = mbar_wait(&conv_bar[s], ph, p.err, 103);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t b_hi = smem_base + s * WH_STAGE_BYTES + WH_RAW_BYTES, b_lo = b_hi + WH_BH_BYTES;
                    const uint32_t a_hi = tmem + WH_ACC_COLS + s * WH_A_COLS, a_lo = a_hi + WH_BK / 2;
                    const uint32_t off1 = (uint32_t)blk0 * 32 * 64;               // half 1's rows follow half 0's
#pragma unroll
                    for (int kk = 0; kk < WH_BK / 16; ++kk) {
                        {
                            const uint64_t dbh = umma_desc_k_sw64(b_hi + kk * 32), dbl = umma_desc_k_sw64(b_lo + kk * 32);
                            umma_f16_ts_2sm(tmem, a_lo + kk * 8, dbh, idesc0, (it | kk) != 0);
                            umma_f16_ts_2sm(tmem, a_hi + kk * 8, dbl, idesc0, 1);
                            umma_f16_ts_2sm(tmem, a_hi + kk * 8, dbh, idesc0, 1);
                        }
                        if (p.h1) {
                            const uint64_t dbh = umma_desc_k_sw64(b_hi + off1 + kk * 32);
                            const uint64_t dbl = umma_desc_k_sw64(b_lo + off1 + kk * 32);
                            const uint32_t d = tmem + p.h0;
                            umma_f16_ts_2sm(d, a_lo + kk * 8, dbh, idesc1, (it | kk) != 0);
                            umma_f16_ts_2sm(d, a_hi + kk * 8, dbl, idesc1, 1);
                            umma_f16_ts_2sm(d, a_hi + kk * 8, dbh, idesc1, 1);
                        }
                    }
                    umma_commit_2sm(&empty_bar[s]);
                    if (it + 1 == total) umma_commit_2sm(&tmem_full_bar);
                }
                __syncwarp();
            }
        }
    } else if (warp < 10) {
        // ------------------------------------------------ A: shifted X rows * s -> TMEM (fp16 pairs); then the epilogue --
        // warp pair (q, half): TMEM lane quarter q = warp % 4, positions [16 half, +16) of the chunk
        const int q = warp & 3, half = (warp - 2) >> 2;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const int grow = m_tile * 256 + (int)rank * 128 + q * 32 + lane;      // output row (tap, n)
        const bool row_ok = grow < p.rows;
        const int wrow = m_tile * 256 + (int)rank * 128 + q * 32;             // the warp's first row: one tap per warp (Nx % 32 == 0)
        const int tap = wrow < p.rows ? wrow / p.Nx : 0;
        const int shift = (tap - p.taps / 2) * p.dilation;
        const float sx = f16_scale_of(__ldg(hp.x_amax));
        uint8_t* a_ring = smem + WH_STAGES * WH_STAGE_BYTES;
        bool ok = true;
        for (int it = 0; it < total && ok; ++it) {
            const int s = it % WH_STAGES, sa = it % WH_ASTAGES;
            const uint32_t ph = (it / WH_STAGES) & 1, pha = (it / WH_ASTAGES) & 1;
            // which of the chunk's 32 positions stay inside their sample after the tap shift (else: the conv's zero padding)
            const int p0 = (it_begin + it) * WH_BK;
            int tj = p0 % p.T + lane;                                    // T >= 32: at most one wrap inside a chunk
            if (tj >= p.T) tj -= p.T;
            const uint32_t inside = __ballot_sync(0xffffffffu, tj + shift >= 0 && tj + shift < p.T);
            const uint32_t live = (row_ok ? inside : 0u) >> (half * 16);      // bit j = this half's position j
            ok = mbar_wait(&afull_bar[sa], pha, p.err, 107);              // this chunk's X boxes have landed
            const float* col = reinterpret_cast<const float*>(a_ring + sa * WH_ATILE_BYTES + q * WP_BLK_BYTES) + lane +
                               half * 16 * 32;
            uint32_t r[16];                                              // [0,8) hi pairs, [8,16) lo pairs of this half
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const float v0 = ((live >> j) & 1u) ? col[j * 32] * sx : 0.f;
                const float v1 = ((live >> (j + 1)) & 1u) ? col[(j + 1) * 32] * sx : 0.f;
                f16_split2(v0, v1, r[j / 2], r[8 + j / 2]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&aempty_bar[sa]);                 // the ring slot can be refilled
            ok = ok && mbar_wait(&empty_bar[s], ph ^ 1, p.err, 104);      // the MMAs of chunk it-STAGES have left this slot
            tc_fence_after();
            tmem_st8u(tq + WH_ACC_COLS + s * WH_A_COLS + half * 8, r);
            tmem_st8u(tq + WH_ACC_COLS + s * WH_A_COLS + WH_BK / 2 + half * 8, r + 8);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&conv_bar[s]), 0));
        }
        // ---- epilogue: partial tile -> workspace [ks][mtiles*256][Mdy]; the two warps of a quarter alternate column chunks ----
        if (!skip) {
            if (total > 0) mbar_wait(&tmem_full_bar, 0, p.err, 106);
            tc_fence_after();
            float* dst = p.P + (((long long)ksl * p.mtiles * 256) + grow) * p.Mdy + n0;
            const float comp = acc_trunc_comp(total * (WH_BK / 16) * 3) * (1.0f / sx) *
                               (1.0f / f16_scale_of(__ldg(hp.dy_amax)));
#pragma unroll 1
            for (int c = half; c < p.nt / 32; c += 2) {
                float v[32];
                if (total > 0) {
                    tmem_ld32(tq + c * 32, v);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] *= comp;
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
            tc_fence_before();
        }
    } else if (warp < 20) {
        // ------------------------------------------------ B: dY block w (32 channels) -> K-major fp16 hi / lo rows ------
        // warp pair (w, half): block w, positions [16 half, +16) = 16-byte chunks 2 half, 2 half + 1 of the 64-byte rows
        const int w = (warp - 10) >> 1, half = (warp - 10) & 1;
        if (w < nblk) {
            const float sdy = f16_scale_of(__ldg(hp.dy_amax));
            const int brow = w * 32 + lane;                              // row of the CTA's B tile (half 0 rows, then half 1 rows)
            const uint32_t sw = (uint32_t)((brow >> 1) & 3);             // SWIZZLE_64B: 16-byte chunk ^= (row / 2) % 4
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % WH_STAGES;
                const uint32_t ph = (it / WH_STAGES) & 1;
                ok = mbar_wait(&full_bar[s], ph, p.err, 105);
                uint8_t* st = smem + s * WH_STAGE_BYTES;
                const float* col = reinterpret_cast<const float*>(st + w * WP_BLK_BYTES) + lane;
                uint8_t* hrow = st + WH_RAW_BYTES + brow * 64;
                uint8_t* lrow = hrow + WH_BH_BYTES;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {                         // 8 positions = one 16-byte chunk of the row
                    const int c = 2 * half + cc;
                    uint4 h, l;
                    f16_split2(col[(8 * c + 0) * 32] * sdy, col[(8 * c + 1) * 32] * sdy, h.x, l.x);
                    f16_split2(col[(8 * c + 2) * 32] * sdy, col[(8 * c + 3) * 32] * sdy, h.y, l.y);
                    f16_split2(col[(8 * c + 4) * 32] * sdy, col[(8 * c + 5) * 32] * sdy, h.z, l.z);
                    f16_split2(col[(8 * c + 6) * 32] * sdy, col[(8 * c + 7) * 32] * sdy, h.w, l.w);
                    *reinterpret_cast<uint4*>(hrow + (((uint32_t)c ^ sw) << 4)) = h;
                    *reinterpret_cast<uint4*>(lrow + (((uint32_t)c ^ sw) << 4)) = l;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(mapa_u32(smem_u32(&conv_bar[s]), 0));
            }
        }
    } else {
        // ------------------------------------------------ TMA producer of the X ring: 4 boxes (one per converter warp) ---
        if (lane == 0) {
            prefetch_tmap(&tmX);
            uint8_t* a_ring = smem + WH_STAGES * WH_STAGE_BYTES;
            int nq[4], sh[4];
#pragma unroll
            for (int w = 0; w < 4; ++w) {
                const int wrow = m_tile * 256 + (int)rank * 128 + w * 32;   // first (tap, n) row of converter quarter w
                const bool okw = wrow < p.rows;                             // Nx % 32 == 0: a quarter never straddles a tap
                const int tapw = okw ? wrow / p.Nx : 0;
                nq[w] = okw ? wrow - tapw * p.Nx : -1;
                sh[w] = (tapw - p.taps / 2) * p.dilation;
            }
            for (int it = 0; it < total; ++it) {
                const int sa = it % WH_ASTAGES;
                const uint32_t pha = (it / WH_ASTAGES) & 1;
                if (!mbar_wait(&aempty_bar[sa], pha ^ 1, p.err, 108)) break;
                const int p0 = (it_begin + it) * WH_BK;
                uint8_t* dst = a_ring + sa * WH_ATILE_BYTES;
                mbar_expect_tx(&afull_bar[sa], WH_ATILE_BYTES);
#pragma unroll
                for (int w = 0; w < 4; ++w)        // padding quarters read past the end of the tensor: zero-filled
                    tma_load_2d(dst + w * WP_BLK_BYTES, &tmX, &afull_bar[sa], nq[w] < 0 ? 0 : nq[w], nq[w] < 0 ? p.R : p0 + sh[w]);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem);
    }
}

// dY [B,T,Mdy], X [B,T,Nx] channels-last -> dW [Mdy][Ntrue][taps]; ws: wgradp_workspace_floats() floats
inline int launch_wgrad_hp(const float* dY, const float* dy_amax, const float* X, const float* x_amax, int B, int T, int Mdy,
                           int Nx, int Ntrue, int taps, int dilation, float* ws, float* dW, int* err, cudaStream_t st) {
    const WgradPPGeom g = wgradp_geometry(B, T, Mdy, Nx, taps);
    const long long R = (long long)B * T;
    if (R >= (1ll << 31)) return set_error(2, "%s: too many rows%s", __func__);
    CUtensorMap tmDY, tmX;
    {
        uint64_t dims[2] = {(uint64_t)Mdy, (uint64_t)R};
        uint64_t str[1] = {(uint64_t)Mdy * 4};
        uint32_t box[2] = {32, 32};
        if (!make_tmap_f32(&tmDY, dY, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return set_error(4, "%s: cuTensorMapEncodeTiled failed%s", __func__);
    }
    {
        uint64_t dims[2] = {(uint64_t)Nx, (uint64_t)R};
        uint64_t str[1] = {(uint64_t)Nx * 4};
        uint32_t box[2] = {32, 32};
        if (!make_tmap_f32(&tmX, X, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_NONE))
            return set_error(4, "%s: cuTensorMapEncodeTiled(X) failed%s", __func__);
    }
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(wgrad_hp_kernel), WH_SMEM_BYTES)) return rc;
    WgradHP hp;
    WgradPP& p = hp.c;
    p.R = (int)R; p.T = T; p.Mdy = Mdy; p.Nx = Nx; p.taps = taps; p.dilation = dilation; p.rows = g.rows;
    p.mtiles = g.mtiles; p.ntiles = g.ntiles; p.ks = g.ks; p.nt = g.nt; p.h0 = g.h0; p.h1 = g.h1;
    p.chunks = g.chunks; p.per_split = g.per_split; p.P = ws; p.err = err; p.dbg = nullptr;
    hp.x_amax = x_amax; hp.dy_amax = dy_amax;
    wgrad_hp_kernel<<<2 * g.mtiles * g.ntiles * g.ks, WH_THREADS, WH_SMEM_BYTES, st>>>(tmDY, tmX, hp);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e));
    wgradp_reduce_kernel<<<ew_grid((long long)Mdy * Ntrue * taps), 256, 0, st>>>(ws, dW, taps, g.ks, g.mtiles * 256, Mdy, Nx,
                                                                                Ntrue);
    ++g_launches;
    e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: reduce launch failed: %s", __func__, cudaGetErrorString(e));
    return 0;
}

}  // namespace tc
}  // namespace bm
