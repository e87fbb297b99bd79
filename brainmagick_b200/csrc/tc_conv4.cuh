// K3/K4/K5, fourth-generation tcgen05 kernel: PERSISTENT CTA pairs with double-buffered accumulators.
//
// Measured on the third generation (tc_conv3.cuh, one 128x320 tile per CTA pair, 768 CTAs per K3 launch): the K loop
// itself ran at 83 % of the MMA rate, but per tile 9 us of epilogue (TMEM -> global) and 3 us of prologue were exposed
// because the 320 accumulator columns + the A staging fill tensor memory, and 5.19 waves of tiles cost 6.
// This generation keeps the CTA-pair MMA (cta_group::2, M = 256, A through tensor memory) but
//   * works on HALF-width tiles (128 positions x NH <= 160 channels per CTA) with TWO accumulator buffers in TMEM, so the
//     epilogue of tile i (4 dedicated warps) overlaps the K loop of tile i+1;
//   * is persistent: each CTA pair walks tiles pair, pair+P, pair+2P, ... without draining its TMA / TMEM pipelines, so
//     there is no per-tile prologue and the tail is half as coarse;
//   * has a 6-deep shared-memory ring (36 KB stages) decoupled from the 3-deep TMEM ring of split activations.
//
// Warp roles (320 threads, both CTAs of the pair): warp 0 TMA producer, warp 1 MMA issuer (leader CTA) + TMEM alloc,
// warps 2-5 split x rows into tf32 hi/lo in TMEM, warps 6-9 epilogue.
// TMEM (512 columns): [0,NH) and [NH,2NH) accumulators, [320 + 64 s, +64) activation stage s = 32 hi + 32 lo.
#pragma once
#include "tc_common.cuh"
#include "tc_conv3.cuh"

namespace bm {
namespace tc {

constexpr int C4_BM = 128, C4_BK = 32, C4_STAGES = 6, C4_TSTAGES = 3, C4_THREADS = 320;
constexpr int C4_MAX_NH = 160;
constexpr int C4_A_BYTES = C4_BM * C4_BK * 4;                        // 16 KB
constexpr int C4_BQ_BYTES_MAX = (C4_MAX_NH / 2) * C4_BK * 4;         // 10 KB: this CTA's half of the weight tile
constexpr int C4_STAGE_BYTES = C4_A_BYTES + 2 * C4_BQ_BYTES_MAX;     // 36 KB (x tile, weights hi, weights lo)
constexpr int C4_SMEM_BYTES = C4_STAGES * C4_STAGE_BYTES + 1024;
constexpr int C4_ACC_COLS = 2 * C4_MAX_NH;                           // 320
constexpr int C4_A_COLS = 2 * C4_BK;                                 // 64 per TMEM stage

struct Conv4P {
    int B, T, Cin, Ntot;
    int taps, dilation, sign;
    int glu, nh, act, out_tmajor;
    int ntiles_n, mtiles, bpairs, total_tiles;
    const float* bias;
    const float* addend;
    float* y;
    float* aux;
    float* glu_out;
    int* err;
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C4_THREADS, 1)
conv_tc4_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                const __grid_constant__ CUtensorMap tmBlo, const Conv4P p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[C4_STAGES], empty_bar[C4_STAGES];
    __shared__ __align__(8) uint64_t conv_bar[C4_TSTAGES], tempty_bar[C4_TSTAGES];
    __shared__ __align__(8) uint64_t acc_full_bar[2], acc_empty_bar[2];
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

    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    const int kchunks = p.Cin / C4_BK;
    const int per_tile = p.taps * kchunks;
    const int H = p.Ntot / 2;
    const int nh = p.nh, nq = nh / 2;
    const uint32_t bq_bytes = (uint32_t)(nq * C4_BK * 4);
    const int my_tiles = skip ? 0 : (p.total_tiles > pair ? (p.total_tiles - pair + npairs - 1) / npairs : 0);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C4_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < C4_TSTAGES; ++s) { mbar_init(&conv_bar[s], 8); mbar_init(&tempty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full_bar[a], 1); mbar_init(&acc_empty_bar[a], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    // tile id -> (n_tile fastest, then m tile, then sample pair)
    auto decode = [&](int ti, int& n_tile, int& t0, int& b) {
        const int tile = pair + ti * npairs;
        n_tile = tile % p.ntiles_n;
        const int r = tile / p.ntiles_n;
        t0 = (r % p.mtiles) * C4_BM;
        b = 2 * (r / p.mtiles) + (int)rank;
    };

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmBhi);
            prefetch_tmap(&tmBlo);
            int g = 0;
            bool ok = true;
            for (int ti = 0; ti < my_tiles && ok; ++ti) {
                int n_tile, t0, b;
                decode(ti, n_tile, t0, b);
                const int rowbase = p.glu ? ((int)rank * H + n_tile * nq) : (n_tile * nh + (int)rank * nq);
                for (int it = 0; it < per_tile; ++it, ++g) {
                    const int s = g % C4_STAGES;
                    const uint32_t ph = (g / C4_STAGES) & 1;
                    ok = mbar_wait(&empty_bar[s], ph ^ 1, p.err, 41);
                    if (!ok) break;
                    const int tap = it / kchunks, k0 = (it - tap * kchunks) * C4_BK;
                    const int shift = p.sign * (tap - p.taps / 2) * p.dilation;
                    uint8_t* st = smem + s * C4_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], C4_A_BYTES + 2 * bq_bytes);
                    tma_load_3d(st, &tmA, &full_bar[s], k0, t0 + shift, b);
                    tma_load_2d(st + C4_A_BYTES, &tmBhi, &full_bar[s], k0, tap * p.Ntot + rowbase);
                    tma_load_2d(st + C4_A_BYTES + C4_BQ_BYTES_MAX, &tmBlo, &full_bar[s], k0, tap * p.Ntot + rowbase);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(2 * C4_BM, nh);
            int g = 0;
            bool ok = true;
            for (int ti = 0; ti < my_tiles && ok; ++ti) {
                const int acc = ti & 1;
                ok = mbar_wait(&acc_empty_bar[acc], ((ti >> 1) & 1) ^ 1, p.err, 42);   // epilogue drained this buffer
                if (!ok) break;
                tc_fence_after();
                const uint32_t d = tmem + acc * nh;
                for (int it = 0; it < per_tile; ++it, ++g) {
                    const int s = g % C4_STAGES, ts = g % C4_TSTAGES;
                    ok = mbar_wait(&conv_bar[ts], (g / C4_TSTAGES) & 1, p.err, 43);
                    if (!ok) break;
                    tc_fence_after();
                    const uint32_t b_hi = smem_base + s * C4_STAGE_BYTES + C4_A_BYTES, b_lo = b_hi + C4_BQ_BYTES_MAX;
                    const uint32_t a_hi = tmem + C4_ACC_COLS + ts * C4_A_COLS, a_lo = a_hi + C4_BK;
#pragma unroll
                    for (int kk = 0; kk < C4_BK / 8; ++kk) {
                        const uint64_t dbh = umma_desc_k_sw128(b_hi + kk * 32), dbl = umma_desc_k_sw128(b_lo + kk * 32);
                        umma_tf32_ts_2sm(d, a_lo + kk * 8, dbh, idesc, (it | kk) != 0);
                        umma_tf32_ts_2sm(d, a_hi + kk * 8, dbl, idesc, 1);
                        umma_tf32_ts_2sm(d, a_hi + kk * 8, dbh, idesc, 1);
                    }
                    umma_commit_2sm(&empty_bar[s]);
                    umma_commit_2sm(&tempty_bar[ts]);
                }
                if (ok) umma_commit_2sm(&acc_full_bar[acc]);
            }
        }
    } else if (warp < 6) {
        // ---- warps 2..5: x rows -> tf32 hi/lo -> tensor memory ----
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t conv_leader[C4_TSTAGES] = {mapa_u32(smem_u32(&conv_bar[0]), 0), mapa_u32(smem_u32(&conv_bar[1]), 0),
                                                  mapa_u32(smem_u32(&conv_bar[2]), 0)};
        const int total = my_tiles * per_tile;
        bool ok = true;
        for (int g = 0; g < total && ok; ++g) {
            const int s = g % C4_STAGES, ts = g % C4_TSTAGES;
            ok = mbar_wait(&full_bar[s], (g / C4_STAGES) & 1, p.err, 44);
            ok = mbar_wait(&tempty_bar[ts], ((g / C4_TSTAGES) & 1) ^ 1, p.err, 45) && ok;
            tc_fence_after();
            const uint8_t* arow = smem + s * C4_STAGE_BYTES + row * 128;
            float hi[C4_BK], lo[C4_BK];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                tf32_split(v.x, hi[4 * c + 0], lo[4 * c + 0]); tf32_split(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                tf32_split(v.z, hi[4 * c + 2], lo[4 * c + 2]); tf32_split(v.w, hi[4 * c + 3], lo[4 * c + 3]);
            }
            tmem_st32(tq + C4_ACC_COLS + ts * C4_A_COLS, hi);
            tmem_st32(tq + C4_ACC_COLS + ts * C4_A_COLS + C4_BK, lo);
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(conv_leader[ts]);
        }
    } else {
        // ---- warps 6..9: epilogue of tile ti while the K loop of tile ti+1 runs ----
        const int q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const uint32_t empty_leader[2] = {mapa_u32(smem_u32(&acc_empty_bar[0]), 0), mapa_u32(smem_u32(&acc_empty_bar[1]), 0)};
        bool ok = true;
        for (int ti = 0; ti < my_tiles && ok; ++ti) {
            const int acc = ti & 1;
            int n_tile, t0, b;
            decode(ti, n_tile, t0, b);
            ok = mbar_wait(&acc_full_bar[acc], (ti >> 1) & 1, p.err, 46);
            tc_fence_after();
            const uint32_t tacc = tq + acc * nh;
            const int t = t0 + row;
            const bool valid = t < p.T && b < p.B;
            if (!p.glu) {
                const int n0 = n_tile * nh;
                const long long off = ((long long)b * p.T + t) * p.Ntot + n0;
#pragma unroll 1
                for (int c = 0; c < nh / 16; ++c) {
                    float v[16];
                    tmem_ld16(tacc + c * 16, v);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                            if (p.bias) {
                                float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + c * 16 + j);
                                o.x += bb.x; o.y += bb.y; o.z += bb.z; o.w += bb.w;
                            }
                            if (p.addend) {
                                float4 aa = *reinterpret_cast<const float4*>(p.addend + off + c * 16 + j);
                                o.x += aa.x; o.y += aa.y; o.z += aa.z; o.w += aa.w;
                            }
                            if (p.aux) *reinterpret_cast<float4*>(p.aux + off + c * 16 + j) = o;
                            if (p.act) { o.x = gelu_f(o.x); o.y = gelu_f(o.y); o.z = gelu_f(o.z); o.w = gelu_f(o.w); }
                            if (!p.out_tmajor) {
                                *reinterpret_cast<float4*>(p.y + off + c * 16 + j) = o;
                            } else {
                                float* yt = p.y + ((long long)b * p.Ntot + n0 + c * 16 + j) * p.T + t;
                                yt[0] = o.x; yt[p.T] = o.y; yt[2 * (long long)p.T] = o.z; yt[3 * (long long)p.T] = o.w;
                            }
                        }
                    }
                }
            } else {
                const int c0 = n_tile * nq;                  // 'a' columns [0,nq), gate columns [nq, nh) of the tile
                const long long rowi = (long long)b * p.T + t;
#pragma unroll 1
                for (int c = 0; c < nq / 16; ++c) {
                    float a[16], gt[16];
                    tmem_ld16(tacc + c * 16, a);
                    tmem_ld16(tacc + nq + c * 16, gt);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 av = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
                            float4 gv = make_float4(gt[j], gt[j + 1], gt[j + 2], gt[j + 3]);
                            if (p.bias) {
                                float4 ba = *reinterpret_cast<const float4*>(p.bias + c0 + c * 16 + j);
                                float4 bg = *reinterpret_cast<const float4*>(p.bias + H + c0 + c * 16 + j);
                                av.x += ba.x; av.y += ba.y; av.z += ba.z; av.w += ba.w;
                                gv.x += bg.x; gv.y += bg.y; gv.z += bg.z; gv.w += bg.w;
                            }
                            if (p.y) {
                                *reinterpret_cast<float4*>(p.y + rowi * p.Ntot + c0 + c * 16 + j) = av;
                                *reinterpret_cast<float4*>(p.y + rowi * p.Ntot + H + c0 + c * 16 + j) = gv;
                            }
                            float4 o = make_float4(av.x * sigmoid_f(gv.x), av.y * sigmoid_f(gv.y),
                                                   av.z * sigmoid_f(gv.z), av.w * sigmoid_f(gv.w));
                            *reinterpret_cast<float4*>(p.glu_out + rowi * H + c0 + c * 16 + j) = o;
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(empty_leader[acc]);      // this warp's quarter of the buffer is drained
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem);
    }
}

// N tile: non-GLU: nh | Ntot; GLU: (nh/2) | H.  nh a multiple of 32 in [64, 160].
inline int conv_tc4_pick_nh(int Ntot, int glu) {
    if (glu && (Ntot % 2)) return 0;
    for (int nh = 160; nh >= 64; nh -= 32) {
        if (!glu && Ntot % nh == 0) return nh;
        if (glu && (Ntot / 2) % (nh / 2) == 0) return nh;
    }
    return 0;
}
inline bool conv_tc4_supported(int T, int Cin, int Ntot, int Kw, int glu) {
    if (Kw < 1 || Kw > 3 || (Kw & 1) == 0) return false;
    if (Cin % C4_BK != 0) return false;
    return conv_tc4_pick_nh(Ntot, glu) != 0;
}

// x [B,T,Cin]; w_hi / w_lo [Kw][Ntot][Cin] (tf32-split K-major weights)
inline int launch_conv_tc4(const float* x, const float* w_hi, const float* w_lo, Conv4P p, cudaStream_t st) {
    p.nh = conv_tc4_pick_nh(p.Ntot, p.glu);
    if (p.nh == 0) return set_error(2, "%s: unsupported N%s", __func__);
    CUtensorMap tmA, tmBh, tmBl;
    {
        uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.Cin * 4, (uint64_t)p.T * p.Cin * 4};
        uint32_t box[3] = {C4_BK, C4_BM, 1};
        if (!make_tmap_f32(&tmA, x, 3, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(A) failed%s", __func__);
    }
    {
        uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)p.taps * p.Ntot};
        uint64_t str[1] = {(uint64_t)p.Cin * 4};
        uint32_t box[2] = {C4_BK, (uint32_t)(p.nh / 2)};
        if (!make_tmap_f32(&tmBh, w_hi, 2, dims, str, box) || !make_tmap_f32(&tmBl, w_lo, 2, dims, str, box))
            return set_error(4, "%s: cuTensorMapEncodeTiled(B) failed%s", __func__);
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C4_SMEM_BYTES);
        if (e != cudaSuccess) return set_error(3, "%s: cudaFuncSetAttribute: %s", __func__, cudaGetErrorString(e));
        attr_set = true;
    }
    p.ntiles_n = p.glu ? (p.Ntot / 2) / (p.nh / 2) : p.Ntot / p.nh;
    p.mtiles = (p.T + C4_BM - 1) / C4_BM;
    p.bpairs = (p.B + 1) / 2;
    p.total_tiles = p.ntiles_n * p.mtiles * p.bpairs;
    int pairs = num_sms() / 2;
    if (pairs > p.total_tiles) pairs = p.total_tiles;
    if (pairs < 1) pairs = 1;
    conv_tc4_kernel<<<dim3(2 * pairs), C4_THREADS, C4_SMEM_BYTES, st>>>(tmA, tmBh, tmBl, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e));
    return 0;
}

}  // namespace tc
}  // namespace bm
