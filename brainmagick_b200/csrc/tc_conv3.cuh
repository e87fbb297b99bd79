// K3/K4/K5, third-generation tcgen05 kernel: CTA PAIRS (cta_group::2).
//
// ncu on the single-CTA kernels showed the real ceiling of 3xTF32 on one SM: SHARED-MEMORY BANDWIDTH.  fp32 operands
// are fat, every weight byte is read three times by the MMAs (hi*hi, lo*hi, hi*lo), and the in-kernel hi/lo split
// adds its own smem traffic: ~160 B/cycle/SM wanted vs 128 available (tensor pipe 41-55 % busy).
// A CTA pair halves the B-operand side: one `tcgen05.mma.cta_group::2` (M = 256) multiplies the 128 rows held by EACH
// CTA against a weight tile of which each CTA stores only HALF the rows, so per SM the tensor core reads 32 B/cycle of
// B instead of 64, TMA writes half as many weight bytes, and L2->SM traffic drops to 29 B/cycle.
//
//   cluster = 2 CTAs = the same 128-position block of two consecutive samples (b = 2*pair + rank);
//   per CTA per K chunk (32):   x tile 16 KB (3-D TMA, OOB rows zero = conv padding)  -> converter warps -> TMEM (hi|lo)
//                               weight rows [rank*NH/2, +NH/2) of both column halves, hi and lo (pre-split), 4 x 10 KB
//   leader CTA, one thread:     24 x tcgen05.mma.cta_group::2.kind::tf32 (M=256, N=NH, K=8, A from TMEM, B from smem)
//   tcgen05.commit multicast releases the stage in both CTAs; each CTA's epilogue drains its own TMEM half.
// Hand-shake: converters of BOTH CTAs arrive (remote mbarrier arrive via mapa) on the LEADER's conv barrier once their
// CTA's x rows are in TMEM and their CTA's weight half has landed, so the MMA thread waits on one barrier only.
#pragma once
#include "tc_common.cuh"

namespace bm {
namespace tc {

constexpr int C3_BM = 128, C3_BK = 32, C3_STAGES = 3, C3_THREADS = 320;
constexpr int C3_MAX_NH = 160;
constexpr int C3_A_BYTES = C3_BM * C3_BK * 4;                        // 16 KB
constexpr int C3_BQ_BYTES_MAX = (C3_MAX_NH / 2) * C3_BK * 4;         // 10 KB: one (half-N, hi or lo) quarter
constexpr int C3_STAGE_BYTES = C3_A_BYTES + 4 * C3_BQ_BYTES_MAX;     // 56 KB
constexpr int C3_SMEM_BYTES = C3_STAGES * C3_STAGE_BYTES + 1024;
constexpr int C3_ACC_COLS = 2 * C3_MAX_NH;                           // 320
constexpr int C3_A_COLS = 2 * C3_BK;                                 // 64 per stage -> 320 + 3*64 = 512

struct Conv3P {
    int B, T, Cin, Ntot;
    int taps, dilation, sign;
    int glu, nh, act, out_tmajor;
    const float* bias;
    const float* addend;
    float* y;
    float* aux;
    float* glu_out;
    int* err;
    int tma_epi;              // 0: per-thread stores; 1: TMA store of 32x32 blocks; 2: TMA reduce-add (y += tile, in place)
    double* stats;            // [2*Ntot] or null: BatchNorm batch statistics sum(y), sum(y^2) accumulated from the staged
                              // epilogue blocks (fp64 atomics; only with tma_epi == 1; zeroed by the launcher)
};

// smem (128B-swizzled 32x32 fp32 block) -> global through the TMA: full-line, asynchronous stores; rows/samples outside
// the tensor are clipped by the hardware.
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// address of the same shared-memory object in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
    return r;
}
// Arrive on an mbarrier of another CTA of the cluster.  Default semantics (.release.cta), NOT .release.cluster: ptxas lowers
// a cluster-scope release to MEMBAR.ALL.GPU + ERRBAR, which cost the converter warps ~1 300 cycles per K chunk (measured
// with the cycle counters of tc_wgradp.cuh: the MMA thread waited for the converters 53 % of the time).  What the consumer
// reads is ordered by other means: tensor memory by the tcgen05.fence::before/after_thread_sync pair around this arrive,
// shared memory written through the generic proxy by the fence.proxy.async before it (the pattern of CUTLASS's
// ClusterBarrier::arrive(cta_id)).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(NCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_2sm(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this smem offset in BOTH CTAs of the pair when the issued MMAs are complete
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
        ::"r"(smem_u32(bar)), "h"((uint16_t)3)
        : "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(C3_THREADS, 1)
conv_tc3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
                const __grid_constant__ CUtensorMap tmBlo, const __grid_constant__ CUtensorMap tmY,
                const __grid_constant__ CUtensorMap tmO, const Conv3P p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[C3_STAGES], conv_bar[C3_STAGES], empty_bar[C3_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int prior_error;

    // (uniform over the grid in practice: the flag is only ever set by a timed-out wait of an earlier launch/CTA)
    if (threadIdx.x == 0) prior_error = p.err ? *reinterpret_cast<volatile int*>(p.err) : 0;
    __syncthreads();
    const bool skip = prior_error != 0;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int n_tile = blockIdx.y, mtiles = (p.T + C3_BM - 1) / C3_BM;
    const int pair = blockIdx.x >> 1;
    const int mt = pair % mtiles;
    const int b = 2 * (pair / mtiles) + (int)rank;        // b >= B for the odd tail: loads zero-fill, stores skipped
    const int t0 = mt * C3_BM;
    const int kchunks = p.Cin / C3_BK;
    const int total = skip ? 0 : p.taps * kchunks;
    const int H = p.Ntot / 2;
    const int nh = p.nh, nq = nh / 2;                      // rows of each column half held by this CTA
    const int rowbase0 = (p.glu ? n_tile * nh : n_tile * 2 * nh) + (int)rank * nq;
    const int rowbase1 = (p.glu ? H + n_tile * nh : n_tile * 2 * nh + nh) + (int)rank * nq;
    const uint32_t bq_bytes = (uint32_t)(nq * C3_BK * 4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C3_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&conv_bar[s], 256);                  // 128 converter threads of each CTA (leader's copy is used)
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc_2sm<512>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                    // both CTAs' barriers + TMEM exist before anyone signals
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmBhi);
            prefetch_tmap(&tmBlo);
            for (int it = 0; it < total; ++it) {
                const int s = it % C3_STAGES;
                const uint32_t ph = (it / C3_STAGES) & 1;
                if (!mbar_wait(&empty_bar[s], ph ^ 1, p.err, 31)) break;
                const int tap = it / kchunks, k0 = (it - tap * kchunks) * C3_BK;
                const int shift = p.sign * (tap - p.taps / 2) * p.dilation;
                uint8_t* st = smem + s * C3_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], C3_A_BYTES + 4 * bq_bytes);
                tma_load_3d(st, &tmA, &full_bar[s], k0, t0 + shift, b);
                uint8_t* bq = st + C3_A_BYTES;
                tma_load_2d(bq, &tmBhi, &full_bar[s], k0, tap * p.Ntot + rowbase0);
                tma_load_2d(bq + C3_BQ_BYTES_MAX, &tmBhi, &full_bar[s], k0, tap * p.Ntot + rowbase1);
                tma_load_2d(bq + 2 * C3_BQ_BYTES_MAX, &tmBlo, &full_bar[s], k0, tap * p.Ntot + rowbase0);
                tma_load_2d(bq + 3 * C3_BQ_BYTES_MAX, &tmBlo, &full_bar[s], k0, tap * p.Ntot + rowbase1);
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(2 * C3_BM, nh);
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % C3_STAGES;
                const uint32_t ph = (it / C3_STAGES) & 1;
                ok = mbar_wait(&conv_bar[s], ph, p.err, 33);
                if (!ok) break;
                tc_fence_after();
                const uint32_t bq = smem_base + s * C3_STAGE_BYTES + C3_A_BYTES;
                const uint32_t a_hi = tmem + C3_ACC_COLS + s * C3_A_COLS, a_lo = a_hi + C3_BK;
#pragma unroll
                for (int kk = 0; kk < C3_BK / 8; ++kk) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t dbh = umma_desc_k_sw128(bq + half * C3_BQ_BYTES_MAX + kk * 32);
                        const uint64_t dbl = umma_desc_k_sw128(bq + (2 + half) * C3_BQ_BYTES_MAX + kk * 32);
                        const uint32_t d = tmem + half * nh;
                        umma_tf32_ts_2sm(d, a_lo + kk * 8, dbh, idesc, (it | kk) != 0);
                        umma_tf32_ts_2sm(d, a_hi + kk * 8, dbl, idesc, 1);
                        umma_tf32_ts_2sm(d, a_hi + kk * 8, dbh, idesc, 1);
                    }
                }
                umma_commit_2sm(&empty_bar[s]);
            }
            umma_commit_2sm(&tmem_full_bar);
        }
    } else {
        const int cw = warp - 2;                         // 0..7
        const int q = warp & 3;
        const bool a_warp = cw < 4;
        const int row = q * 32 + lane;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        if (a_warp) {
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % C3_STAGES;
                const uint32_t ph = (it / C3_STAGES) & 1;
                ok = mbar_wait(&full_bar[s], ph, p.err, 34);   // this CTA's x tile AND weight quarter-tiles landed
                const uint8_t* arow = smem + s * C3_STAGE_BYTES + row * 128;
                float hi[C3_BK], lo[C3_BK];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                    tf32_split(v.x, hi[4 * c + 0], lo[4 * c + 0]); tf32_split(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                    tf32_split(v.z, hi[4 * c + 2], lo[4 * c + 2]); tf32_split(v.w, hi[4 * c + 3], lo[4 * c + 3]);
                }
                tmem_st32(tq + C3_ACC_COLS + s * C3_A_COLS, hi);
                tmem_st32(tq + C3_ACC_COLS + s * C3_A_COLS + C3_BK, lo);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive_cluster(mapa_u32(smem_u32(&conv_bar[s]), 0));   // the LEADER's barrier, from either CTA
            }
        }
        // ---------------- epilogue: two warps per TMEM lane quarter, each takes half of the columns ----------------
        if (!skip) {
            mbar_wait(&tmem_full_bar, 0, p.err, 35);
            tc_fence_after();
            const int cset = cw >> 2;
            const int t = t0 + row;
            const bool valid = t < p.T && b < p.B;
            if (p.tma_epi == 3) {
                // GLU through the TMA: per 32-column chunk three staged 32x32 blocks -- h's `a` part, h's gate part (both
                // only when the pre-activation is saved for backward) and out = a * sigmoid(gate)
                const int nch = nh / 32;
                const int ch_begin = cset == 0 ? 0 : (nch + 1) / 2, ch_end = cset == 0 ? (nch + 1) / 2 : nch;
                const int c0 = n_tile * nh;
                uint8_t* stg = smem + cw * 12288;
                if (lane == 0) { prefetch_tmap(&tmY); prefetch_tmap(&tmO); }
#pragma unroll 1
                for (int c = ch_begin; c < ch_end; ++c) {
                    float a[32], g[32];
                    tmem_ld32(tq + c * 32, a);
                    tmem_ld32(tq + nh + c * 32, g);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 ba = *reinterpret_cast<const float4*>(p.bias + c0 + c * 32 + j);
                            const float4 bg = *reinterpret_cast<const float4*>(p.bias + H + c0 + c * 32 + j);
                            a[j] += ba.x; a[j + 1] += ba.y; a[j + 2] += ba.z; a[j + 3] += ba.w;
                            g[j] += bg.x; g[j + 1] += bg.y; g[j + 2] += bg.z; g[j + 3] += bg.w;
                        }
                    }
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    uint8_t* rowp = stg + lane * 128;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int sw = (j ^ (lane & 7)) << 4;
                        if (p.y) {
                            *reinterpret_cast<float4*>(rowp + sw) = make_float4(a[4 * j], a[4 * j + 1], a[4 * j + 2], a[4 * j + 3]);
                            *reinterpret_cast<float4*>(rowp + 4096 + sw) =
                                make_float4(g[4 * j], g[4 * j + 1], g[4 * j + 2], g[4 * j + 3]);
                        }
                        *reinterpret_cast<float4*>(rowp + 8192 + sw) =
                            make_float4(a[4 * j] * sigmoid_f(g[4 * j]), a[4 * j + 1] * sigmoid_f(g[4 * j + 1]),
                                        a[4 * j + 2] * sigmoid_f(g[4 * j + 2]), a[4 * j + 3] * sigmoid_f(g[4 * j + 3]));
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.y) {
                            tma_store_3d(&tmY, stg, c0 + c * 32, t0 + q * 32, b);
                            tma_store_3d(&tmY, stg + 4096, H + c0 + c * 32, t0 + q * 32, b);
                        }
                        tma_store_3d(&tmO, stg + 8192, c0 + c * 32, t0 + q * 32, b);
                        bulk_commit();
                    }
                }
                if (lane == 0) bulk_wait<0>();
                __syncwarp();
            } else if (p.tma_epi) {
                // Measured: the per-thread 16-byte stores below (1280 B apart) made the epilogue 18 % of the kernel.  Here each
                // warp writes its 32x32 accumulator block into a 128B-swizzled shared-memory tile (the pipeline stages are
                // free once tmem_full has fired in both CTAs) and one lane hands it to the TMA: full-line asynchronous
                // stores, or a reduce-add straight into the skip-path gradient (y += tile) with no addend loads at all.
                const int ncol0 = cset * nh;
                const int n0 = n_tile * 2 * nh + ncol0;
                uint8_t* stg = smem + cw * 8192;                       // two 4 KB buffers per warp
                if (lane == 0) prefetch_tmap(&tmY);
#pragma unroll 1
                for (int c = 0; c < nh / 32; ++c) {
                    float v[32];
                    tmem_ld32(tq + ncol0 + c * 32, v);
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bb = *reinterpret_cast<const float4*>(p.bias + n0 + c * 32 + j);
                            v[j] += bb.x; v[j + 1] += bb.y; v[j + 2] += bb.z; v[j + 3] += bb.w;
                        }
                    }
                    uint8_t* buf = stg + (c & 1) * 4096;
                    if (lane == 0) bulk_wait_read<1>();                // the store that used this buffer two chunks ago
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(buf + lane * 128 + ((j ^ (lane & 7)) << 4)) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if (p.tma_epi == 2) tma_reduce_add_3d(&tmY, buf, n0 + c * 32, t0 + q * 32, b);
                        else tma_store_3d(&tmY, buf, n0 + c * 32, t0 + q * 32, b);
                        bulk_commit();
                    }
                    if (p.stats) {
                        // column sums of this 32x32 block straight from the staged tile (lane = column): the separate
                        // statistics pass over y (118 MB per layer) disappears
                        const int nrows = (b < p.B) ? min(32, p.T - (t0 + q * 32)) : 0;
                        float s1 = 0.f, s2 = 0.f;
                        for (int r = 0; r < nrows; ++r) {
                            const float x = *reinterpret_cast<const float*>(buf + r * 128 + (((lane >> 2) ^ (r & 7)) << 4) +
                                                                              (lane & 3) * 4);
                            s1 += x;
                            s2 = fmaf(x, x, s2);
                        }
                        if (nrows > 0) {
                            atomicAdd(p.stats + n0 + c * 32 + lane, (double)s1);
                            atomicAdd(p.stats + p.Ntot + n0 + c * 32 + lane, (double)s2);
                        }
                    }
                }
                if (lane == 0) bulk_wait<0>();
                __syncwarp();
            } else if (!p.glu) {
                const int ncol0 = cset * nh;
                const int n0 = n_tile * 2 * nh + ncol0;
                const long long off = ((long long)b * p.T + t) * p.Ntot + n0;
#pragma unroll 1
                for (int c = 0; c < nh / 16; ++c) {
                    float v[16];
                    tmem_ld16(tq + ncol0 + c * 16, v);
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
                const int j0 = cset * (nh / 2);
                const int c0 = n_tile * nh + j0;
                const long long rowi = (long long)b * p.T + t;
#pragma unroll 1
                for (int c = 0; c < (nh / 2) / 16; ++c) {
                    float a[16], g[16];
                    tmem_ld16(tq + j0 + c * 16, a);
                    tmem_ld16(tq + nh + j0 + c * 16, g);
                    if (valid) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            float4 av = make_float4(a[j], a[j + 1], a[j + 2], a[j + 3]);
                            float4 gv = make_float4(g[j], g[j + 1], g[j + 2], g[j + 3]);
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
        }
    }
    __syncthreads();
    cluster_sync_all();                                    // nobody frees TMEM / exits while the pair is still in flight
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2sm<512>(tmem);
    }
}

inline int conv_tc3_pick_nh(int Ntot, int glu) {
    if (glu && (Ntot % 2)) return 0;
    const int n = glu ? Ntot / 2 : Ntot;
    for (int nh = 160; nh >= 128; nh -= 32) {
        const int per_tile = glu ? nh : 2 * nh;
        if (n % per_tile == 0) return nh;
    }
    return 0;
}
inline bool conv_tc3_supported(int T, int Cin, int Ntot, int Kw, int glu) {
    if (Kw < 1 || Kw > 3 || (Kw & 1) == 0) return false;
    if (Cin % C3_BK != 0) return false;
    return conv_tc3_pick_nh(Ntot, glu) != 0;
}

// x [B,T,Cin]; w_hi / w_lo [Kw][Ntot][Cin] (tf32-split K-major weights)
inline int launch_conv_tc3(const float* x, const float* w_hi, const float* w_lo, Conv3P p, cudaStream_t st) {
    p.nh = conv_tc3_pick_nh(p.Ntot, p.glu);
    if (p.nh == 0) return set_error(2, "%s: unsupported N%s", __func__);
    CUtensorMap tmA, tmBh, tmBl, tmY, tmO;
    {
        uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.Cin * 4, (uint64_t)p.T * p.Cin * 4};
        uint32_t box[3] = {C3_BK, C3_BM, 1};
        if (!make_tmap_f32(&tmA, x, 3, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(A) failed%s", __func__);
    }
    {
        uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)p.taps * p.Ntot};
        uint64_t str[1] = {(uint64_t)p.Cin * 4};
        uint32_t box[2] = {C3_BK, (uint32_t)(p.nh / 2)};
        if (!make_tmap_f32(&tmBh, w_hi, 2, dims, str, box) || !make_tmap_f32(&tmBl, w_lo, 2, dims, str, box))
            return set_error(4, "%s: cuTensorMapEncodeTiled(B) failed%s", __func__);
    }
    // TMA epilogue for the plain modes: store, or in-place accumulate when the caller passes addend == y
    p.tma_epi = 0;
    if (!p.glu && !p.act && !p.aux && !p.out_tmajor && p.y && (p.nh % 32 == 0) && (p.addend == nullptr || p.addend == p.y))
        p.tma_epi = p.addend ? 2 : 1;
    {
        uint64_t dims[3] = {(uint64_t)p.Ntot, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.Ntot * 4, (uint64_t)p.T * p.Ntot * 4};
        uint32_t box[3] = {32, 32, 1};
        const void* base = p.tma_epi ? (const void*)p.y : (const void*)x;       // any valid tensor when unused
        if (!p.tma_epi) { dims[0] = (uint64_t)p.Cin; str[0] = (uint64_t)p.Cin * 4; str[1] = (uint64_t)p.T * p.Cin * 4; }
        if (!make_tmap_f32(&tmY, base, 3, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(Y) failed%s", __func__);
        tmO = tmY;
    }
    if (p.glu && p.nh % 32 == 0 && p.glu_out && !p.stats) {
        // GLU through the TMA: tmY describes h [B,T,2H] (when saved), tmO the gated output [B,T,H]
        p.tma_epi = 3;
        const int Hh = p.Ntot / 2;
        uint32_t box[3] = {32, 32, 1};
        uint64_t dimo[3] = {(uint64_t)Hh, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t stro[2] = {(uint64_t)Hh * 4, (uint64_t)p.T * Hh * 4};
        if (!make_tmap_f32(&tmO, p.glu_out, 3, dimo, stro, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(O) failed%s", __func__);
        if (p.y) {
            uint64_t dimh[3] = {(uint64_t)p.Ntot, (uint64_t)p.T, (uint64_t)p.B};
            uint64_t strh[2] = {(uint64_t)p.Ntot * 4, (uint64_t)p.T * p.Ntot * 4};
            if (!make_tmap_f32(&tmY, p.y, 3, dimh, strh, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(H) failed%s", __func__);
        }
    }
    if (int rc_ = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc3_kernel), C3_SMEM_BYTES)) return rc_;
    const int ntiles = p.glu ? (p.Ntot / 2) / p.nh : p.Ntot / (2 * p.nh);
    const int mtiles = (p.T + C3_BM - 1) / C3_BM;
    const int pairs = mtiles * ((p.B + 1) / 2);
    dim3 grid(2 * pairs, ntiles, 1);
    conv_tc3_kernel<<<grid, C3_THREADS, C3_SMEM_BYTES, st>>>(tmA, tmBh, tmBl, tmY, tmO, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e));
    return 0;
}

}  // namespace tc
}  // namespace bm
