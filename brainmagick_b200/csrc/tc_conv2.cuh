// K3/K4/K5, second-generation tcgen05 kernel: one CTA = 128 positions x (2 x NH) output channels, A through TMEM.
//
// Why (ncu, profiles/r1b_ncu_conv_tc_summary.csv): the first kernel (tc_conv.cuh, 128x160 tile, both operands and both tf32
// halves of the weights streamed from L2 into shared memory) is L2->SM bandwidth bound: 56 KB of operands per 960 MMA
// cycles per SM, tensor pipe 41 % busy.  This version
//   * computes BOTH column halves of a position tile in one CTA (x tile fetched once per K chunk, not twice);
//   * fetches the weights as raw fp32 and splits them into tf32 hi/lo in shared memory (no second weight stream);
//   * moves the activation operand out of shared memory altogether: the converter warps read each x row from the
//     TMA-written (128B-swizzled) stage, split it, and write it to TENSOR MEMORY (tcgen05.st); the MMA is the
//     "TS" form (A from TMEM, B from smem), which also halves the tensor core's shared-memory read traffic.
// => 56 KB fresh bytes per 1920 MMA cycles per SM (29 B/cycle), smem operand reads 62 B/cycle.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2-9 converters + epilogue.
// TMEM (512 columns): [0, 2*NH) two fp32 accumulators, [320 + 64 s, +64) A stage s = 32 hi + 32 lo columns.
#pragma once
#include "tc_common.cuh"

namespace bm {
namespace tc {

constexpr int C2_BM = 128, C2_BK = 32, C2_STAGES = 2, C2_THREADS = 320, C2_CONV_THREADS = 256;
constexpr int C2_MAX_NH = 160;
constexpr int C2_A_BYTES = C2_BM * C2_BK * 4;                    // 16 KB
constexpr int C2_B_BYTES_MAX = 2 * C2_MAX_NH * C2_BK * 4;        // 40 KB
constexpr int C2_STAGE_BYTES = C2_A_BYTES + 2 * C2_B_BYTES_MAX;  // 96 KB
constexpr int C2_SMEM_BYTES = C2_STAGES * C2_STAGE_BYTES + 1024;
constexpr int C2_ACC_COLS = 2 * C2_MAX_NH;                       // 320
constexpr int C2_A_COLS = 2 * C2_BK;                             // 64 per stage

struct Conv2P {
    int B, T, Cin, Ntot;
    int taps, dilation, sign;
    int glu;                  // halves = 'a' columns [c0, c0+nh) and gate columns [H+c0, H+c0+nh)
    int nh;                   // 160 or 128
    int act, out_tmajor;
    int ksplit;               // >1: split the K chunks over gridDim.z = B*ksplit CTAs; y receives [ksplit][B][T][Ntot]
                              //     partial sums (no bias / activation / addend), reduced by the caller
    const float* bias;
    const float* addend;
    float* y;
    float* aux;
    float* glu_out;
    int* err;
};

__global__ void __launch_bounds__(C2_THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const Conv2P p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[C2_STAGES], conv_bar[C2_STAGES], empty_bar[C2_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int prior_error;

    if (threadIdx.x == 0) prior_error = p.err ? *reinterpret_cast<volatile int*>(p.err) : 0;
    __syncthreads();
    if (prior_error) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int n_tile = blockIdx.x, t0 = blockIdx.y * C2_BM;
    const int b = blockIdx.z / p.ksplit, split = blockIdx.z - b * p.ksplit;
    const int kchunks = p.Cin / C2_BK;
    const int all_chunks = p.taps * kchunks;
    const int per_split = (all_chunks + p.ksplit - 1) / p.ksplit;
    const int it_begin = split * per_split;
    const int total = max(0, min(all_chunks, it_begin + per_split) - it_begin);     // this CTA's K chunks
    const int H = p.Ntot / 2;
    const int nh = p.nh;
    // weight rows (within one tap) of the two column halves of this tile
    const int rowbase0 = p.glu ? n_tile * nh : n_tile * 2 * nh;
    const int rowbase1 = p.glu ? H + n_tile * nh : n_tile * 2 * nh + nh;
    const uint32_t b_bytes = (uint32_t)(2 * nh * C2_BK * 4);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C2_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&conv_bar[s], C2_CONV_THREADS);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmB);
            for (int it = 0; it < total; ++it) {
                const int s = it % C2_STAGES;
                const uint32_t ph = (it / C2_STAGES) & 1;
                if (!mbar_wait(&empty_bar[s], ph ^ 1, p.err, 21)) break;
                const int git = it_begin + it;
                const int tap = git / kchunks, k0 = (git - tap * kchunks) * C2_BK;
                const int shift = p.sign * (tap - p.taps / 2) * p.dilation;
                uint8_t* st = smem + s * C2_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], C2_A_BYTES + b_bytes);
                tma_load_3d(st, &tmA, &full_bar[s], k0, t0 + shift, b);
                tma_load_2d(st + C2_A_BYTES, &tmB, &full_bar[s], k0, tap * p.Ntot + rowbase0);
                tma_load_2d(st + C2_A_BYTES + nh * C2_BK * 4, &tmB, &full_bar[s], k0, tap * p.Ntot + rowbase1);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(C2_BM, nh);
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % C2_STAGES;
                const uint32_t ph = (it / C2_STAGES) & 1;
                ok = mbar_wait(&conv_bar[s], ph, p.err, 23);
                if (!ok) break;
                tc_fence_after();
                const uint32_t b_hi = smem_base + s * C2_STAGE_BYTES + C2_A_BYTES, b_lo = b_hi + C2_B_BYTES_MAX;
                const uint32_t a_hi = tmem + C2_ACC_COLS + s * C2_A_COLS, a_lo = a_hi + C2_BK;
#pragma unroll
                for (int kk = 0; kk < C2_BK / 8; ++kk) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t boff = half * nh * C2_BK * 4 + kk * 32;
                        const uint64_t dbh = umma_desc_k_sw128(b_hi + boff), dbl = umma_desc_k_sw128(b_lo + boff);
                        const uint32_t d = tmem + half * nh;
                        umma_tf32_ts(d, a_lo + kk * 8, dbh, idesc, (it | kk) != 0);
                        umma_tf32_ts(d, a_hi + kk * 8, dbl, idesc, 1);
                        umma_tf32_ts(d, a_hi + kk * 8, dbh, idesc, 1);
                    }
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        const int cw = warp - 2;                         // 0..7
        const int q = warp & 3;                          // TMEM lane quarter
        const int ct = cw * 32 + lane;                   // 0..255
        const bool a_warp = cw < 4;                      // warps 2..5 also stage the A rows
        const int row = q * 32 + lane;                   // tile row this thread owns (A staging and epilogue)
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        bool ok = true;
        for (int it = 0; it < total && ok; ++it) {
            const int s = it % C2_STAGES;
            const uint32_t ph = (it / C2_STAGES) & 1;
            ok = mbar_wait(&full_bar[s], ph, p.err, 24);
            uint8_t* st = smem + s * C2_STAGE_BYTES;
            if (a_warp) {
                // x row -> tf32 hi/lo -> tensor memory (lane = row, 32 columns each)
                const uint8_t* arow = st + row * 128;
                float hi[C2_BK], lo[C2_BK];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(arow + ((c ^ (row & 7)) << 4));
                    tf32_split(v.x, hi[4 * c + 0], lo[4 * c + 0]); tf32_split(v.y, hi[4 * c + 1], lo[4 * c + 1]);
                    tf32_split(v.z, hi[4 * c + 2], lo[4 * c + 2]); tf32_split(v.w, hi[4 * c + 3], lo[4 * c + 3]);
                }
                tmem_st32(tq + C2_ACC_COLS + s * C2_A_COLS, hi);
                tmem_st32(tq + C2_ACC_COLS + s * C2_A_COLS + C2_BK, lo);
                tmem_st_wait();
            }
            // weights: raw fp32 -> hi (in place) and lo
            float4* bh = reinterpret_cast<float4*>(st + C2_A_BYTES);
            float4* bl = reinterpret_cast<float4*>(st + C2_A_BYTES + C2_B_BYTES_MAX);
            const int nvec = (int)(b_bytes / 16);
#pragma unroll
            for (int j = 0; j < C2_B_BYTES_MAX / 16 / C2_CONV_THREADS; ++j) {
                const int idx = ct + C2_CONV_THREADS * j;
                if (idx < nvec) {
                    float4 v = bh[idx], h, l;
                    tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
                    bh[idx] = h;
                    bl[idx] = l;
                }
            }
            fence_proxy_async();
            tc_fence_before();
            mbar_arrive(&conv_bar[s]);
        }
        // ---------------- epilogue: two warps per TMEM lane quarter, each takes half of the columns ----------------
        mbar_wait(&tmem_full_bar, 0, p.err, 25);
        tc_fence_after();
        const int cset = cw >> 2;                        // 0: warps 2-5, 1: warps 6-9
        const int t = t0 + row;
        const bool valid = t < p.T;
        if (!p.glu) {
            const int ncol0 = cset * nh;                 // this warp's accumulator columns [ncol0, ncol0 + nh)
            const int n0 = n_tile * 2 * nh + ncol0;
            const long long off = (((long long)split * p.B + b) * p.T + t) * p.Ntot + n0;
#pragma unroll 1
            for (int c = 0; c < nh / 16; ++c) {
                float v[16];
                tmem_ld16(tq + ncol0 + c * 16, v);
                if (total == 0) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) v[j] = 0.f;
                }
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
            const int j0 = cset * (nh / 2);              // this warp's 'a' columns [j0, j0 + nh/2), gates at nh + same
            const int c0 = n_tile * nh + j0;             // first output channel
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
                        float4 o = make_float4(av.x * sigmoid_f(gv.x), av.y * sigmoid_f(gv.y), av.z * sigmoid_f(gv.z),
                                               av.w * sigmoid_f(gv.w));
                        *reinterpret_cast<float4*>(p.glu_out + rowi * H + c0 + c * 16 + j) = o;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

inline int conv_tc2_pick_nh(int Ntot, int glu) {
    const int n = glu ? Ntot / 2 : Ntot;                 // columns covered by n-tiles of (glu ? nh : 2*nh)
    if (glu && (Ntot % 2)) return 0;
    for (int nh = 160; nh >= 128; nh -= 32) {
        const int per_tile = glu ? nh : 2 * nh;
        if (n % per_tile == 0) return nh;
    }
    return 0;
}
inline bool conv_tc2_supported(int T, int Cin, int Ntot, int Kw, int glu) {
    if (Kw < 1 || Kw > 3 || (Kw & 1) == 0) return false;
    if (Cin % C2_BK != 0) return false;
    return conv_tc2_pick_nh(Ntot, glu) != 0;
}

// x [B,T,Cin]; w_raw [Kw][Ntot][Cin] (fp32, K-major re-layout of the nn.Conv1d weight)
inline int launch_conv_tc2(const float* x, const float* w_raw, Conv2P p, cudaStream_t st) {
    p.nh = conv_tc2_pick_nh(p.Ntot, p.glu);
    if (p.nh == 0) return set_error(2, "%s: unsupported N%s", __func__);
    CUtensorMap tmA, tmB;
    {
        uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.Cin * 4, (uint64_t)p.T * p.Cin * 4};
        uint32_t box[3] = {C2_BK, C2_BM, 1};
        if (!make_tmap_f32(&tmA, x, 3, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(A) failed%s", __func__);
    }
    {
        uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)p.taps * p.Ntot};
        uint64_t str[1] = {(uint64_t)p.Cin * 4};
        uint32_t box[2] = {C2_BK, (uint32_t)p.nh};
        if (!make_tmap_f32(&tmB, w_raw, 2, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(B) failed%s", __func__);
    }
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C2_SMEM_BYTES);
        if (e != cudaSuccess) return set_error(3, "%s: cudaFuncSetAttribute: %s", __func__, cudaGetErrorString(e));
        attr_set = true;
    }
    const int ntiles = p.glu ? (p.Ntot / 2) / p.nh : p.Ntot / (2 * p.nh);
    if (p.ksplit < 1) p.ksplit = 1;
    if (p.ksplit > 1 && (p.glu || p.bias || p.addend || p.act || p.aux || p.out_tmajor))
        return set_error(2, "%s: split-K supports plain outputs only%s", __func__);
    dim3 grid(ntiles, (p.T + C2_BM - 1) / C2_BM, p.B * p.ksplit);
    conv_tc2_kernel<<<grid, C2_THREADS, C2_SMEM_BYTES, st>>>(tmA, tmB, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e));
    return 0;
}

}  // namespace tc
}  // namespace bm

namespace bm {
namespace tc {
// out[r][c] = colscale[c] * sum_s P[s][r][c]   (fixed order => deterministic)
__global__ void splitk_reduce_scale_kernel(const float* __restrict__ P, const float* __restrict__ colscale,
                                           float* __restrict__ out, int ksplit, long long rows, int cols) {
    long long total = rows * cols;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < ksplit; ++k) s += P[(long long)k * total + i];
        out[i] = colscale ? s * colscale[(int)(i % cols)] : s;
    }
}
}  // namespace tc
}  // namespace bm
