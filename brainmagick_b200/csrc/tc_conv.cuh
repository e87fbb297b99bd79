// K3/K4 on the 5th-generation tensor cores: dilated Conv1d as an implicit GEMM with 3xTF32 error compensation.
//
//   y[b, t, n] = bias[n] + sum_{tap} sum_{k} x[b, t + shift(tap), k] * W[tap][n][k]      (channels-last x, y)
//
// Roles inside one CTA (192 threads), one 128(t) x 160(n) output tile per CTA:
//   warp 0    TMA producer: per (tap, 32-wide k chunk) one 3-D box of x (rows shifted by the tap; rows outside
//             [0,T) are zero-filled by TMA = the conv's zero padding) and the matching rows of W_hi / W_lo,
//             all in the 128-byte-swizzled K-major layout tcgen05 reads.
//   warps 2-5 split the fp32 x tile in shared memory into tf32 hi and lo parts (in place + a second buffer),
//             then act as the epilogue: TMEM -> registers -> bias / skip-gradient add / GLU -> global.
//   warp 1    one thread issues tcgen05.mma (kind::tf32, M=128, N=160, K=8): hi*hi + lo*hi + hi*lo into an
//             fp32 accumulator in tensor memory; tcgen05.commit releases the stage back to the producer.
// Every mbarrier wait is bounded (tc_common.cuh) so a pipeline bug reports an error instead of hanging the GPU.
#pragma once
#include "tc_common.cuh"

namespace bm {
namespace tc {

constexpr int CV_BM = 128, CV_BN = 160, CV_BK = 32, CV_STAGES = 3, CV_THREADS = 192;
constexpr int CV_A_BYTES = CV_BM * CV_BK * 4;       // 16 KB
constexpr int CV_B_BYTES = CV_BN * CV_BK * 4;       // 20 KB
constexpr int CV_STAGE_BYTES = 2 * CV_A_BYTES + 2 * CV_B_BYTES;
constexpr int CV_SMEM_BYTES = CV_STAGES * CV_STAGE_BYTES + 1024;

struct ConvTcP {
    int B, T, Cin, Ntot;      // Ntot = output channels of the GEMM (2H for the GLU conv)
    int taps, dilation, sign;
    int glu;                  // 1: tile = 80 'a' columns + 80 gate columns, out = a * sigmoid(gate)
    int bn;                   // N tile (multiple of 16, <= 160; 160 when glu)
    int act;                  // 1: exact GELU on the output (after bias/addend); `aux` receives the pre-activation
    int out_tmajor;           // 1: y is [B, Ntot, T] (channel-major, the encoder's output layout)
    const float* bias;        // [Ntot] or null
    const float* addend;      // [B,T,Ntot] or null
    float* y;                 // [B,T,Ntot] (pre-GLU h when glu; may be null then)
    float* aux;               // [B,T,Ntot] or null
    float* glu_out;           // [B,T,Ntot/2]
    int n_wsets;              // number of weight sets when wsel != null
    const int* wsel;          // per-sample weight set (SubjectLayers): weight rows are offset by wsel[b]*taps*Ntot; or null
    int* err;
};

__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBhi,
               const __grid_constant__ CUtensorMap tmBlo, const ConvTcP p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[CV_STAGES], conv_bar[CV_STAGES], empty_bar[CV_STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_base_smem;
    __shared__ int prior_error;

    // a previous CTA already reported a pipeline failure: do not spend another timeout on it
    if (threadIdx.x == 0) prior_error = p.err ? *reinterpret_cast<volatile int*>(p.err) : 0;
    __syncthreads();
    if (prior_error) return;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));

    const int n_tile = blockIdx.x, t0 = blockIdx.y * CV_BM, b = blockIdx.z;
    const int kchunks = p.Cin / CV_BK;
    const int total = p.taps * kchunks;
    const int H = p.Ntot / 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < CV_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&conv_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(&tmem_base_smem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmA);
            prefetch_tmap(&tmBhi);
            prefetch_tmap(&tmBlo);
            int wsel = p.wsel ? p.wsel[b] : 0;
            if (wsel < 0 || wsel >= p.n_wsets) {             // a subject / recording index outside the weight sets: the TMA would
                if (p.err) atomicCAS(p.err, 0, 900);         // zero-fill silently; report it like the reference's gather would raise
                wsel = 0;
            }
            const int wsel_base = wsel * p.taps;
            for (int it = 0; it < total; ++it) {
                const int s = it % CV_STAGES;
                const uint32_t ph = (it / CV_STAGES) & 1;
                if (!mbar_wait(&empty_bar[s], ph ^ 1, p.err, 1)) break;
                const int tap = it / kchunks, k0 = (it - tap * kchunks) * CV_BK;
                const int shift = p.sign * (tap - p.taps / 2) * p.dilation;
                uint8_t* st = smem + s * CV_STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], CV_A_BYTES + 2 * p.bn * CV_BK * 4);
                tma_load_3d(st, &tmA, &full_bar[s], k0, t0 + shift, b);
                uint8_t* bh = st + 2 * CV_A_BYTES;
                uint8_t* bl = bh + CV_B_BYTES;
                if (!p.glu) {
                    const int row = (wsel_base + tap) * p.Ntot + n_tile * p.bn;
                    tma_load_2d(bh, &tmBhi, &full_bar[s], k0, row);
                    tma_load_2d(bl, &tmBlo, &full_bar[s], k0, row);
                } else {
                    const int ra = tap * p.Ntot + n_tile * (CV_BN / 2), rg = ra + H;
                    tma_load_2d(bh, &tmBhi, &full_bar[s], k0, ra);
                    tma_load_2d(bh + CV_B_BYTES / 2, &tmBhi, &full_bar[s], k0, rg);
                    tma_load_2d(bl, &tmBlo, &full_bar[s], k0, ra);
                    tma_load_2d(bl + CV_B_BYTES / 2, &tmBlo, &full_bar[s], k0, rg);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(CV_BM, p.bn);
            bool ok = true;
            for (int it = 0; it < total && ok; ++it) {
                const int s = it % CV_STAGES;
                const uint32_t ph = (it / CV_STAGES) & 1;
                ok = mbar_wait(&full_bar[s], ph, p.err, 2) && mbar_wait(&conv_bar[s], ph, p.err, 3);
                if (!ok) break;
                tc_fence_after();
                const uint32_t a_hi = smem_base + s * CV_STAGE_BYTES;
                const uint32_t a_lo = a_hi + CV_A_BYTES;
                const uint32_t b_hi = a_lo + CV_A_BYTES;
                const uint32_t b_lo = b_hi + CV_B_BYTES;
#pragma unroll
                for (int kk = 0; kk < CV_BK / 8; ++kk) {
                    const uint32_t off = kk * 32;      // 8 tf32 = 32 bytes along K inside the 128-byte swizzle span
                    const uint64_t dah = umma_desc_k_sw128(a_hi + off), dal = umma_desc_k_sw128(a_lo + off);
                    const uint64_t dbh = umma_desc_k_sw128(b_hi + off), dbl = umma_desc_k_sw128(b_lo + off);
                    umma_tf32_ss(tmem, dal, dbh, idesc, (it | kk) != 0);
                    umma_tf32_ss(tmem, dah, dbl, idesc, 1);
                    umma_tf32_ss(tmem, dah, dbh, idesc, 1);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // ---- converter: split the x tile into tf32 hi (in place) and lo ----
        const int ct = (warp - 2) * 32 + lane;          // 0..127
        bool ok = true;
        for (int it = 0; it < total && ok; ++it) {
            const int s = it % CV_STAGES;
            const uint32_t ph = (it / CV_STAGES) & 1;
            ok = mbar_wait(&full_bar[s], ph, p.err, 4);
            float4* hi = reinterpret_cast<float4*>(smem + s * CV_STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * CV_STAGE_BYTES + CV_A_BYTES);
#pragma unroll
            for (int j = 0; j < CV_A_BYTES / 16 / 128; ++j) {
                const int idx = ct + 128 * j;
                float4 v = hi[idx], h, l;
                tf32_split(v.x, h.x, l.x); tf32_split(v.y, h.y, l.y); tf32_split(v.z, h.z, l.z); tf32_split(v.w, h.w, l.w);
                hi[idx] = h;
                lo[idx] = l;
            }
            fence_proxy_async();
            mbar_arrive(&conv_bar[s]);
        }
        // ---- epilogue ----
        mbar_wait(&tmem_full_bar, 0, p.err, 5);
        tc_fence_after();
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int t = t0 + q * 32 + lane;
        const bool valid = t < p.T;
        const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16);
        const float comp = acc_trunc_comp(total * (CV_BK / 8) * 3);          // additions chained into the accumulator
        if (!p.glu) {
            const int n0 = n_tile * p.bn;
            const long long off = ((long long)b * p.T + t) * p.Ntot + n0;
#pragma unroll 1
            for (int c = 0; c < p.bn / 16; ++c) {
                float v[16];
                tmem_ld16(tq + c * 16, v);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] *= comp;
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
            const int c0 = n_tile * (CV_BN / 2);
            const long long row = (long long)b * p.T + t;
#pragma unroll 1
            for (int c = 0; c < (CV_BN / 2) / 16; ++c) {
                float a[16], g[16];
                tmem_ld16(tq + c * 16, a);
                tmem_ld16(tq + CV_BN / 2 + c * 16, g);
#pragma unroll
                for (int j = 0; j < 16; ++j) { a[j] *= comp; g[j] *= comp; }
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
                            *reinterpret_cast<float4*>(p.y + row * p.Ntot + c0 + c * 16 + j) = av;
                            *reinterpret_cast<float4*>(p.y + row * p.Ntot + H + c0 + c * 16 + j) = gv;
                        }
                        float4 o = make_float4(av.x * sigmoid_f(gv.x), av.y * sigmoid_f(gv.y), av.z * sigmoid_f(gv.z),
                                               av.w * sigmoid_f(gv.w));
                        *reinterpret_cast<float4*>(p.glu_out + row * H + c0 + c * 16 + j) = o;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<256>(tmem);
    }
}

// W[o][i][j] (nn.Conv1d layout) -> tf32-split K-major operands:
//   fwd  operand: F[j][o][i]   (n = o, k = i)        data-gradient operand: G[j][i][o]   (n = i, k = o)
__global__ void weight_split_kernel(const float* __restrict__ W, float* __restrict__ f_hi, float* __restrict__ f_lo,
                                    float* __restrict__ g_hi, float* __restrict__ g_lo, int O, int I, int Kw) {
    long long total = (long long)O * I * Kw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int j = (int)(idx % Kw);
        long long oi = idx / Kw;
        int i = (int)(oi % I), o = (int)(oi / I);
        float v = W[idx];
        float h = tf32_rna(v), l = tf32_rna(v - h);
        long long fi = ((long long)j * O + o) * I + i, gi = ((long long)j * I + i) * O + o;
        // a NULL lo pointer asks for the RAW fp32 value in the re-laid array (the v2 kernel splits in shared memory)
        if (f_hi) { if (f_lo) { f_hi[fi] = h; f_lo[fi] = l; } else f_hi[fi] = v; }
        if (g_hi) { if (g_lo) { g_hi[gi] = h; g_lo[gi] = l; } else g_hi[gi] = v; }
    }
}

// N tile for a GEMM with Ntot output columns: the largest multiple of 16 <= 160 that divides Ntot (0: unsupported)
inline int conv_tc_pick_bn(int Ntot, int glu) {
    if (glu) return (Ntot % 2 == 0 && (Ntot / 2) % (CV_BN / 2) == 0) ? CV_BN : 0;
    for (int bn = CV_BN; bn >= 64; bn -= 16)
        if (Ntot % bn == 0) return bn;
    return 0;
}
inline bool conv_tc_supported(int T, int Cin, int Ntot, int Kw, int glu) {
    if (Kw < 1 || Kw > 3 || (Kw & 1) == 0) return false;
    if (Cin % CV_BK != 0) return false;
    return conv_tc_pick_bn(Ntot, glu) != 0;
}

// x [B,T,Cin]; w_hi/w_lo [Kw][Ntot][Cin]
inline int launch_conv_tc(const float* x, const float* w_hi, const float* w_lo, ConvTcP p, cudaStream_t st) {
    p.bn = conv_tc_pick_bn(p.Ntot, p.glu);
    if (p.bn == 0) return set_error(2, "%s: unsupported N%s", __func__);
    CUtensorMap tmA, tmBh, tmBl;
    {
        uint64_t dims[3] = {(uint64_t)p.Cin, (uint64_t)p.T, (uint64_t)p.B};
        uint64_t str[2] = {(uint64_t)p.Cin * 4, (uint64_t)p.T * p.Cin * 4};
        uint32_t box[3] = {CV_BK, CV_BM, 1};
        if (!make_tmap_f32(&tmA, x, 3, dims, str, box)) return set_error(4, "%s: cuTensorMapEncodeTiled(A) failed%s", __func__);
    }
    {
        uint64_t dims[2] = {(uint64_t)p.Cin, (uint64_t)p.taps * p.Ntot * (uint64_t)(p.wsel ? p.n_wsets : 1)};
        uint64_t str[1] = {(uint64_t)p.Cin * 4};
        uint32_t box[2] = {CV_BK, (uint32_t)(p.glu ? CV_BN / 2 : p.bn)};
        if (!make_tmap_f32(&tmBh, w_hi, 2, dims, str, box) || !make_tmap_f32(&tmBl, w_lo, 2, dims, str, box))
            return set_error(4, "%s: cuTensorMapEncodeTiled(B) failed%s", __func__);
    }
    if (int rc_ = ensure_dyn_smem(reinterpret_cast<const void*>(conv_tc_kernel), CV_SMEM_BYTES)) return rc_;
    dim3 grid(p.glu ? (p.Ntot / 2) / (CV_BN / 2) : p.Ntot / p.bn, (p.T + CV_BM - 1) / CV_BM, p.B);
    conv_tc_kernel<<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(tmA, tmBh, tmBl, p);
    ++g_launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return set_error(3, "%s: launch failed: %s", __func__, cudaGetErrorString(e));
    return 0;
}

}  // namespace tc
}  // namespace bm
