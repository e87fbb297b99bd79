// HBM-bound kernels of the hot path: everything that is not a contraction.
// Activations are channels-last [B, T, C] inside the encoder (rows = B*T), so a warp always walks
// contiguous channels; per-channel parameters are read through the read-only path.
#pragma once
#include "common.cuh"

namespace bm {

// ------------------------------------------------------------------------------------------------
// K1 pieces: Fourier embedding of sensor positions and the masked softmax (common.py:254-271,339-357)
// ------------------------------------------------------------------------------------------------
// emb[rc][k*n+l] = cos(loc), emb[rc][P/2 + k*n+l] = sin(loc), loc = (x+margin)*f[k] + (y+margin)*f[l]
// `freq` is the table 2*pi*arange(n)/(1+2*margin) built by the host with the reference's own op order, and
// the products/sum are rounded separately (no FMA contraction) so that `loc` is bit-identical to torch's.
__global__ void fourier_emb_kernel(const float* __restrict__ pos, const float* __restrict__ freq, float margin,
                                   int RC, int n, float* __restrict__ emb) {
    const int P2 = n * n;
    long long total = (long long)RC * P2;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int rc = (int)(idx / P2);
        int kl = (int)(idx - (long long)rc * P2);
        int k = kl / n, l = kl - k * n;
        float x = __fadd_rn(pos[2 * rc], margin), y = __fadd_rn(pos[2 * rc + 1], margin);
        float loc = __fadd_rn(__fmul_rn(x, freq[k]), __fmul_rn(y, freq[l]));
        float s, c;
        sincosf(loc, &s, &c);
        emb[(long long)rc * 2 * P2 + kl] = c;
        emb[(long long)rc * 2 * P2 + P2 + kl] = s;
    }
}

// in-place softmax over c of scores[r][o][:] with -inf on invalid (pos == INVALID) and banned sensors
// (||pos - centre|| <= radius, training only).  One warp per (r, o) row.
__global__ void masked_softmax_kernel(float* __restrict__ w, const float* __restrict__ pos,
                                      const float* __restrict__ centre, float radius, float invalid, int R, int O,
                                      int C) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= R * O) return;
    int r = row / O;
    float* s = w + (long long)row * C;
    const float* pr = pos + (long long)r * C * 2;
    float cx = 0.f, cy = 0.f;
    if (centre) { cx = centre[0]; cy = centre[1]; }
    float mx = -INFINITY;
    for (int c = lane; c < C; c += 32) {
        float x = pr[2 * c], y = pr[2 * c + 1];
        bool masked = (x == invalid) && (y == invalid);
        if (centre) {
            float dx = x - cx, dy = y - cy;
            masked = masked || (sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))) <= radius);
        }
        float v = masked ? -INFINITY : s[c];
        s[c] = v;
        mx = fmaxf(mx, v);
    }
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < C; c += 32) {
        float e = expf(s[c] - mx);
        s[c] = e;
        sum += e;
    }
    sum = warp_sum(sum);
    float inv = 1.f / sum;
    for (int c = lane; c < C; c += 32) s[c] *= inv;
}

// dscore = w * (dw - sum_c w*dw)   (softmax backward, one warp per row)
__global__ void softmax_bwd_kernel(const float* __restrict__ w, const float* __restrict__ dw,
                                   float* __restrict__ ds, int rows, int C) {
    int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* wr = w + (long long)row * C;
    const float* dr = dw + (long long)row * C;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot += wr[c] * dr[c];
    dot = warp_sum(dot);
    for (int c = lane; c < C; c += 32) ds[(long long)row * C + c] = wr[c] * (dr[c] - dot);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm1d + GELU + skip  (common.py:118-120,146-147)
// ------------------------------------------------------------------------------------------------
// stats[0:C] = sum y, stats[C:2C] = sum y^2 (fp64) over n = B*T rows  ->  mean, invstd (biased var),
// running stats update with the unbiased variance (momentum), as nn.BatchNorm1d in training mode.
__global__ void bn_finalize_kernel(const double* __restrict__ stats, double n, float eps, float momentum,
                                   float* running_mean, float* running_var, float* __restrict__ mean,
                                   float* __restrict__ invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double m = stats[c] / n;
    double var = stats[C + c] / n - m * m;
    if (var < 0) var = 0;
    mean[c] = (float)m;
    invstd[c] = (float)(1.0 / sqrt(var + (double)eps));
    if (running_mean) {
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
        double unb = n > 1 ? var * n / (n - 1) : var;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
    }
}

// eval mode: mean = running_mean, invstd = 1/sqrt(running_var + eps)
__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps,
                                     float* __restrict__ mean, float* __restrict__ invstd, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    mean[c] = rm[c];
    invstd[c] = 1.f / sqrtf(rv[c] + eps);
}

// x_new = GELU((y - mean) * invstd * gamma + beta) (+ x_old)
template <int VEC>
__global__ void bn_gelu_skip_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean,
                                        const float* __restrict__ invstd, const float* __restrict__ gamma,
                                        const float* __restrict__ beta, const float* __restrict__ x_old,
                                        float* __restrict__ x_new, long long total, int C) {
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * VEC; i < total;
         i += (long long)gridDim.x * blockDim.x * VEC) {
        int c = (int)(i % C);
        float v[VEC], o[VEC];
        if (VEC == 4) {
            float4 t = *reinterpret_cast<const float4*>(y + i);
            v[0] = t.x; v[1 % VEC] = t.y; v[2 % VEC] = t.z; v[3 % VEC] = t.w;
            if (x_old) {
                float4 u = *reinterpret_cast<const float4*>(x_old + i);
                o[0] = u.x; o[1 % VEC] = u.y; o[2 % VEC] = u.z; o[3 % VEC] = u.w;
            }
        } else {
            v[0] = y[i];
            if (x_old) o[0] = x_old[i];
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            float z = (v[j] - mean[c + j]) * invstd[c + j] * gamma[c + j] + beta[c + j];
            float a = gelu_f(z);
            v[j] = x_old ? a + o[j] : a;
        }
        if (VEC == 4) *reinterpret_cast<float4*>(x_new + i) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
        else x_new[i] = v[0];
    }
}

// backward pass 1: sums[c] += sum dz, sums[C+c] += sum dz*yhat  with dz = g * GELU'(z)
__global__ void bn_gelu_bwd_reduce_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                          const float* __restrict__ mean, const float* __restrict__ invstd,
                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                          double* __restrict__ sums, long long rows, int C, int rows_per_block) {
    int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = min(rows, r0 + rows_per_block);
    float mu = mean[c], is = invstd[c], ga = gamma[c], be = beta[c];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        float yh = (y[r * C + c] - mu) * is;
        float z = yh * ga + be;
        float dz = g[r * C + c] * gelu_grad_f(z);
        s1 += dz;
        s2 += dz * yh;
    }
    atomicAdd(sums + c, (double)s1);
    atomicAdd(sums + C + c, (double)s2);
}

// dgamma = sum dz*yhat, dbeta = sum dz  (float outputs from the fp64 sums)
__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta, int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    dbeta[c] = (float)sums[c];
    dgamma[c] = (float)sums[C + c];
}

// backward pass 2: dy = gamma*invstd*(dz - mean(dz) - yhat*mean(dz*yhat))
// eval mode (use_batch_stats = 0): dy = gamma*invstd*dz
__global__ void bn_gelu_bwd_apply_kernel(const float* __restrict__ g, const float* __restrict__ y,
                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                         const float* __restrict__ gamma, const float* __restrict__ beta,
                                         const double* __restrict__ sums, double n, int use_batch_stats,
                                         float* __restrict__ dy, long long total, int C) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        float is = invstd[c], ga = gamma[c];
        float yh = (y[i] - mean[c]) * is;
        float z = yh * ga + beta[c];
        float dz = g[i] * gelu_grad_f(z);
        float v = dz;
        if (use_batch_stats) v = dz - (float)(sums[c] / n) - yh * (float)(sums[C + c] / n);
        dy[i] = ga * is * v;
    }
}

// ------------------------------------------------------------------------------------------------
// GLU backward (common.py:133-138): h = [a | b] per row (2H columns), out = a*sigmoid(b)
// ------------------------------------------------------------------------------------------------
__global__ void glu_bwd_kernel(const float* __restrict__ g, const float* __restrict__ h, float* __restrict__ dh,
                               long long rows, int H) {
    long long total = rows * H;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long r = i / H;
        int c = (int)(i - r * H);
        float a = h[r * 2 * H + c], b = h[r * 2 * H + H + c];
        float s = sigmoid_f(b);
        float gg = g[i];
        dh[r * 2 * H + c] = gg * s;
        dh[r * 2 * H + H + c] = gg * a * s * (1.f - s);
    }
}

// dh = dq * GELU'(h)   (head, simpleconv.py:185-189)
__global__ void gelu_bwd_kernel(const float* __restrict__ dq, const float* __restrict__ h, float* __restrict__ dh,
                                long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        dh[i] = dq[i] * gelu_grad_f(h[i]);
}

// ------------------------------------------------------------------------------------------------
// bias gradients
// ------------------------------------------------------------------------------------------------
// out[c] += sum_r X[r][c]   (channels-last rows); out must be zeroed by the caller
__global__ void colsum_cl_kernel(const float* __restrict__ X, float* __restrict__ out, long long rows, int C,
                                 int rows_per_block) {
    int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) s += X[r * C + c];
    atomicAdd(out + c, s);
}

// out[c] += sum_r X[r*ld + c]   (rows with a leading dimension ld >= C)
__global__ void colsum_strided_kernel(const float* __restrict__ X, float* __restrict__ out, long long rows, int C,
                                      int ld, int rows_per_block) {
    int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) s += X[r * ld + c];
    atomicAdd(out + c, s);
}

// stats[c] += sum_r X[r][c], stats[C+c] += sum_r X[r][c]^2   (fp64 atomics; BatchNorm batch statistics)
__global__ void col_stats_kernel(const float* __restrict__ X, double* __restrict__ stats, long long rows, int C,
                                 int rows_per_block) {
    int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= C) return;
    long long r0 = (long long)blockIdx.x * rows_per_block;
    long long r1 = min(rows, r0 + rows_per_block);
    float s = 0.f, q = 0.f;
#pragma unroll 4
    for (long long r = r0; r < r1; ++r) {
        float v = X[r * C + c];
        s += v;
        q = fmaf(v, v, q);
    }
    atomicAdd(stats + c, (double)s);
    atomicAdd(stats + C + c, (double)q);
}

// out[n] += sum_{z,t} X[z][n][t]  (channel-major [Z,N,T]); one warp per (z, n) row
__global__ void rowsum_cm_kernel(const float* __restrict__ X, float* __restrict__ out, int Z, int N, int T) {
    long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
    int lane = threadIdx.x & 31;
    if (row >= (long long)Z * N) return;
    const float* x = X + row * T;
    float s = 0.f;
    for (int t = lane; t < T; t += 32) s += x[t];
    s = warp_sum(s);
    if (lane == 0) atomicAdd(out + (int)(row % N), s);
}

// ------------------------------------------------------------------------------------------------
// conv weight re-layout: W[o][i][j] -> Wf[j][i][o] (forward B operand) and Wb[j][o][i] (data-gradient B operand)
// ------------------------------------------------------------------------------------------------
__global__ void weight_prep_kernel(const float* __restrict__ W, float* __restrict__ Wf, float* __restrict__ Wb,
                                   int O, int I, int Kw) {
    long long total = (long long)O * I * Kw;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        int j = (int)(idx % Kw);
        long long oi = idx / Kw;
        int i = (int)(oi % I), o = (int)(oi / I);
        float v = W[idx];
        if (Wf) Wf[((long long)j * I + i) * O + o] = v;
        if (Wb) Wb[((long long)j * O + o) * I + i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// ClipLoss pieces (losses.py:91-114)
// ------------------------------------------------------------------------------------------------
// ss[o] += sum_k cand[o][k]^2 (fp64 atomics; grid = (rows, splits))
__global__ void row_sumsq_kernel(const float* __restrict__ X, double* __restrict__ ss, long long K) {
    __shared__ double red[32];
    const float* x = X + (long long)blockIdx.x * K;
    long long chunk = (K + gridDim.y - 1) / gridDim.y;
    long long k0 = (long long)blockIdx.y * chunk, k1 = min(K, k0 + chunk);
    float s = 0.f;
    double acc = 0.0;
    int cnt = 0;
    for (long long k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
        float v = x[k];
        s = fmaf(v, v, s);
        if (++cnt == 64) { acc += (double)s; s = 0.f; cnt = 0; }
    }
    acc += (double)s;
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) atomicAdd(ss + blockIdx.x, v);
    }
}

// inv_norm[o] = 1 / (1e-8 + sqrt(ss[o]))    (losses.py:91)
__global__ void inv_norm_kernel(const double* __restrict__ ss, float* __restrict__ inv, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) inv[i] = 1.f / (1e-8f + (float)sqrt(ss[i]));
}

// per row b: lse, row_loss[b] = lse - s[b][b+off]; optional probs = softmax(s)   (one block per row)
__global__ void clip_ce_rows_kernel(const float* __restrict__ scores, int Bn, int Bc, int target_offset,
                                    float* __restrict__ row_loss, float* __restrict__ probs) {
    __shared__ float red[32];
    __shared__ float bcast;
    int b = blockIdx.x;
    const float* s = scores + (long long)b * Bc;
    float mx = -INFINITY;
    for (int o = threadIdx.x; o < Bc; o += blockDim.x) mx = fmaxf(mx, s[o]);
    mx = warp_max(mx);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : -INFINITY;
        v = warp_max(v);
        if (threadIdx.x == 0) bcast = v;
    }
    __syncthreads();
    mx = bcast;
    __syncthreads();
    float sum = 0.f;
    for (int o = threadIdx.x; o < Bc; o += blockDim.x) sum += expf(s[o] - mx);
    sum = warp_sum(sum);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
        if (threadIdx.x == 0) bcast = v;
    }
    __syncthreads();
    sum = bcast;
    if (row_loss && threadIdx.x == 0) row_loss[b] = (logf(sum) + mx) - s[b + target_offset];
    if (probs) {
        float inv = 1.f / sum;
        for (int o = threadIdx.x; o < Bc; o += blockDim.x) probs[(long long)b * Bc + o] = expf(s[o] - mx) * inv;
    }
}

// loss = mean(row_loss)  (single block, deterministic order)
__global__ void mean_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)x[i];
    s = warp_sum_d(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        v = warp_sum_d(v);
        if (threadIdx.x == 0) out[0] = (float)(v / n);
    }
}

// G[b][o] = gout * (softmax(s)[b][o] - [o == b+off]) / Bn * inv_norm[o]    (A.6 of SURVEY.md)
// transposed != 0: G is written as [Bc][Bn] (the dY operand layout of the tensor-core dE GEMM)
__global__ void clip_ce_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ inv_norm,
                                   const float* __restrict__ gout, int Bn, int Bc, int target_offset,
                                   float* __restrict__ G, int transposed) {
    long long total = (long long)Bn * Bc;
    float gs = gout[0] / (float)Bn;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        int b = (int)(i / Bc), o = (int)(i - (long long)b * Bc);
        float p = probs[i] - (o == b + target_offset ? 1.f : 0.f);
        float v = gs * p * inv_norm[o];
        if (transposed) G[(long long)o * Bn + b] = v;
        else G[i] = v;
    }
}

// y[b][c][t] = x[b][c][t] * mask[c]   (simpleconv.py:200-203: `subsample_meg_channels` zeroes the sensors not drawn)
__global__ void channel_mask_kernel(const float* __restrict__ x, const float* __restrict__ mask, float* __restrict__ y,
                                    int C, int T, long long total) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)((i / T) % C);
        y[i] = x[i] * mask[c];
    }
}

// in [Z][N][T] -> out [Z][T][N]   (32x32 tiles through shared memory, both sides coalesced)
__global__ void transpose_nt_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int T, int ld_out,
                                    int n_store) {
    // n_store = N: plain transpose; n_store = ld_out > N: the pad columns [N, ld_out) are written as zeros
    __shared__ float tile[32][33];
    const long long zoff = (long long)blockIdx.z * N * T;
    const long long zout = (long long)blockIdx.z * T * ld_out;
    int t = blockIdx.x * 32 + threadIdx.x;
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {
        int n = blockIdx.y * 32 + i;
        tile[i][threadIdx.x] = (n < N && t < T) ? in[zoff + (long long)n * T + t] : 0.f;
    }
    __syncthreads();
    int n = blockIdx.y * 32 + threadIdx.x;
#pragma unroll
    for (int i = threadIdx.y; i < 32; i += 8) {
        int tt = blockIdx.x * 32 + i;
        if (n < n_store && tt < T) out[zout + (long long)tt * ld_out + n] = tile[threadIdx.x][i];
    }
}

// ------------------------------------------------------------------------------------------------
// float4 versions (C % 4 == 0, 16-byte aligned rows): one thread owns 4 consecutive channels
// ------------------------------------------------------------------------------------------------
__global__ void bn_gelu_bwd_apply_v4_kernel(const float4* __restrict__ g, const float4* __restrict__ y,
                                            const float4* __restrict__ mean, const float4* __restrict__ invstd,
                                            const float4* __restrict__ gamma, const float4* __restrict__ beta,
                                            const double* __restrict__ sums, double n, int use_batch_stats,
                                            float4* __restrict__ dy, long long total4, int C) {
    const int C4 = C >> 2;
    const float rn = (float)(1.0 / n);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4), c = c4 * 4;
        const float4 mu = mean[c4], is = invstd[c4], ga = gamma[c4], be = beta[c4];
        const float4 yv = y[i], gv = g[i];
        float m1[4] = {0.f, 0.f, 0.f, 0.f}, m2[4] = {0.f, 0.f, 0.f, 0.f};
        if (use_batch_stats) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { m1[j] = (float)(sums[c + j] / n); m2[j] = (float)(sums[C + c + j] / n); }
        }
        (void)rn;
        float4 o;
        float yh, dz;
        yh = (yv.x - mu.x) * is.x; dz = gv.x * gelu_grad_f(yh * ga.x + be.x); o.x = ga.x * is.x * (dz - m1[0] - yh * m2[0]);
        yh = (yv.y - mu.y) * is.y; dz = gv.y * gelu_grad_f(yh * ga.y + be.y); o.y = ga.y * is.y * (dz - m1[1] - yh * m2[1]);
        yh = (yv.z - mu.z) * is.z; dz = gv.z * gelu_grad_f(yh * ga.z + be.z); o.z = ga.z * is.z * (dz - m1[2] - yh * m2[2]);
        yh = (yv.w - mu.w) * is.w; dz = gv.w * gelu_grad_f(yh * ga.w + be.w); o.w = ga.w * is.w * (dz - m1[3] - yh * m2[3]);
        dy[i] = o;
    }
}

__global__ void glu_bwd_v4_kernel(const float4* __restrict__ g, const float4* __restrict__ h, float4* __restrict__ dh,
                                  long long rows, int H) {
    const int H4 = H >> 2;
    long long total = rows * H4;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / H4;
        const int c4 = (int)(i - r * H4);
        const float4 a = h[r * 2 * H4 + c4], b = h[r * 2 * H4 + H4 + c4], gg = g[i];
        float4 da, db;
        float s;
        s = sigmoid_f(b.x); da.x = gg.x * s; db.x = gg.x * a.x * s * (1.f - s);
        s = sigmoid_f(b.y); da.y = gg.y * s; db.y = gg.y * a.y * s * (1.f - s);
        s = sigmoid_f(b.z); da.z = gg.z * s; db.z = gg.z * a.z * s * (1.f - s);
        s = sigmoid_f(b.w); da.w = gg.w * s; db.w = gg.w * a.w * s * (1.f - s);
        dh[r * 2 * H4 + c4] = da;
        dh[r * 2 * H4 + H4 + c4] = db;
    }
}

__global__ void gelu_bwd_v4_kernel(const float4* __restrict__ dq, const float4* __restrict__ h, float4* __restrict__ dh,
                                   long long total4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total4;
         i += (long long)gridDim.x * blockDim.x) {
        const float4 d = dq[i], x = h[i];
        dh[i] = make_float4(d.x * gelu_grad_f(x.x), d.y * gelu_grad_f(x.y), d.z * gelu_grad_f(x.z), d.w * gelu_grad_f(x.w));
    }
}

// ------------------------------------------------------------------------------------------------
// Column-stationary versions of the three BatchNorm + GELU kernels (channels-last rows, C % 4 == 0, C/4 <= 256):
// block = (C/4, RY) threads; thread (cx, ry) owns columns 4cx..4cx+3 for the rows ry, ry+RY, ... of the block's row range, so
// the per-column parameters live in registers (no i % C, no parameter reloads, no fp64 divisions per element) and every
// warp-level access is a contiguous row piece.  Rows are unrolled by 4 for memory-level parallelism.
// ------------------------------------------------------------------------------------------------
constexpr int CS_ROWS_PER_BLOCK = 96, CS_U = 4;       // rows per block; rows per thread whose loads are issued together

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4s(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }   // streamed once

// max |.| of what a block has just written -> ONE atomicMax per block on `cell` (the bits of a non-negative float order like
// unsigned integers).  The consumer is the F16-pipe conv (tc_convh.cuh), which scales its operand by a power of two from it:
// fusing the reduction here saves bm_amax's extra pass over the tensor.  Every thread of the block must call.
__device__ __forceinline__ float amax4(float m, const float4& o) {
    return fmaxf(fmaxf(m, fmaxf(fabsf(o.x), fabsf(o.y))), fmaxf(fabsf(o.z), fabsf(o.w)));
}
__device__ __forceinline__ void block_amax_commit(float m, unsigned int* cell) {
    __shared__ float warp_max[32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    if ((tid & 31) == 0) warp_max[tid >> 5] = m;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < (nthr + 31) / 32; ++w) m = fmaxf(m, warp_max[w]);
        if (m > 0.f) atomicMax(cell, __float_as_uint(m));
    }
}

// x_new = GELU((y - mean) * invstd * gamma + beta) (+ x_old)
__global__ void __launch_bounds__(320, 2)
bn_gelu_skip_fwd_cs_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ invstd,
                           const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ x_old,
                           float* __restrict__ x_new, long long rows, int C, unsigned int* __restrict__ amax) {
    const int c = threadIdx.x * 4, RY = blockDim.y;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c), be = ld4(beta + c);
    const long long r0 = (long long)blockIdx.x * CS_ROWS_PER_BLOCK;
    const long long r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
    float am = 0.f;
    for (long long r = r0 + threadIdx.y; r < r1; r += (long long)CS_U * RY) {
        float4 v[CS_U], u[CS_U];
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) {
                v[k] = ld4s(y + rr * C + c);
                if (x_old) u[k] = ld4(x_old + rr * C + c);
            }
        }
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) {
                float4 o;
                // same operation order as the reference BatchNorm: ((y - mean) * invstd) * gamma + beta
                o.x = gelu_f(fmaf((v[k].x - mu.x) * is.x, ga.x, be.x));
                o.y = gelu_f(fmaf((v[k].y - mu.y) * is.y, ga.y, be.y));
                o.z = gelu_f(fmaf((v[k].z - mu.z) * is.z, ga.z, be.z));
                o.w = gelu_f(fmaf((v[k].w - mu.w) * is.w, ga.w, be.w));
                if (x_old) { o.x += u[k].x; o.y += u[k].y; o.z += u[k].z; o.w += u[k].w; }
                *reinterpret_cast<float4*>(x_new + rr * C + c) = o;
                am = amax4(am, o);
            }
        }
    }
    if (amax) block_amax_commit(am, amax);
}

// backward pass 1: sums[c] += sum dz, sums[C+c] += sum dz*yhat  with dz = g * GELU'(z); fp32 per thread, combined over the
// block's RY row lanes in shared memory, ONE fp64 atomic pair per column per block
__global__ void __launch_bounds__(320, 2)
bn_gelu_bwd_reduce_cs_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ mean,
                             const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                             double* __restrict__ sums, long long rows, int C) {
    extern __shared__ float red[];                       // [RY][2][C]
    const int c = threadIdx.x * 4, RY = blockDim.y;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c), be = ld4(beta + c);
    const long long r0 = (long long)blockIdx.x * CS_ROWS_PER_BLOCK;
    const long long r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
    float4 s1 = make_float4(0.f, 0.f, 0.f, 0.f), s2 = s1;
    for (long long r = r0 + threadIdx.y; r < r1; r += (long long)CS_U * RY) {
        float4 yv[CS_U], gv[CS_U];
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) { yv[k] = ld4(y + rr * C + c); gv[k] = ld4(g + rr * C + c); }
        }
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) {
                float yh, dz;
                yh = (yv[k].x - mu.x) * is.x; dz = gv[k].x * gelu_grad_f(fmaf(yh, ga.x, be.x)); s1.x += dz; s2.x = fmaf(dz, yh, s2.x);
                yh = (yv[k].y - mu.y) * is.y; dz = gv[k].y * gelu_grad_f(fmaf(yh, ga.y, be.y)); s1.y += dz; s2.y = fmaf(dz, yh, s2.y);
                yh = (yv[k].z - mu.z) * is.z; dz = gv[k].z * gelu_grad_f(fmaf(yh, ga.z, be.z)); s1.z += dz; s2.z = fmaf(dz, yh, s2.z);
                yh = (yv[k].w - mu.w) * is.w; dz = gv[k].w * gelu_grad_f(fmaf(yh, ga.w, be.w)); s1.w += dz; s2.w = fmaf(dz, yh, s2.w);
            }
        }
    }
    float* mine = red + (size_t)threadIdx.y * 2 * C;
    *reinterpret_cast<float4*>(mine + c) = s1;
    *reinterpret_cast<float4*>(mine + C + c) = s2;
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int i = tid; i < 2 * C; i += nthr) {
        double acc = 0.0;
        for (int ry = 0; ry < RY; ++ry) acc += (double)red[(size_t)ry * 2 * C + i];
        atomicAdd(sums + i, acc);
    }
}

// backward pass 2: dy = gamma*invstd*(dz - mean(dz) - yhat*mean(dz*yhat)); the two means come as dbeta/n and dgamma/n
// (written by bn_param_grad_kernel from the fp64 sums); eval mode (use_batch_stats = 0): dy = gamma*invstd*dz
__global__ void __launch_bounds__(320, 2)
bn_gelu_bwd_apply_cs_kernel(const float* __restrict__ g, const float* __restrict__ y, const float* __restrict__ mean,
                            const float* __restrict__ invstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                            const float* __restrict__ dgamma, const float* __restrict__ dbeta, float rn, int use_batch_stats,
                            float* __restrict__ dy, long long rows, int C, unsigned int* __restrict__ amax) {
    const int c = threadIdx.x * 4, RY = blockDim.y;
    const float4 mu = ld4(mean + c), is = ld4(invstd + c), ga = ld4(gamma + c), be = ld4(beta + c);
    float am = 0.f;
    float4 m1 = make_float4(0.f, 0.f, 0.f, 0.f), m2 = m1;
    if (use_batch_stats) {
        const float4 db = ld4(dbeta + c), dg = ld4(dgamma + c);
        m1 = make_float4(db.x * rn, db.y * rn, db.z * rn, db.w * rn);
        m2 = make_float4(dg.x * rn, dg.y * rn, dg.z * rn, dg.w * rn);
    }
    const float4 kk = make_float4(ga.x * is.x, ga.y * is.y, ga.z * is.z, ga.w * is.w);
    const long long r0 = (long long)blockIdx.x * CS_ROWS_PER_BLOCK;
    const long long r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
    for (long long r = r0 + threadIdx.y; r < r1; r += (long long)CS_U * RY) {
        float4 yv[CS_U], gv[CS_U];
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) { yv[k] = ld4s(y + rr * C + c); gv[k] = ld4s(g + rr * C + c); }
        }
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) {
                float4 o;
                float yh, dz;
                yh = (yv[k].x - mu.x) * is.x; dz = gv[k].x * gelu_grad_f(fmaf(yh, ga.x, be.x)); o.x = kk.x * (dz - m1.x - yh * m2.x);
                yh = (yv[k].y - mu.y) * is.y; dz = gv[k].y * gelu_grad_f(fmaf(yh, ga.y, be.y)); o.y = kk.y * (dz - m1.y - yh * m2.y);
                yh = (yv[k].z - mu.z) * is.z; dz = gv[k].z * gelu_grad_f(fmaf(yh, ga.z, be.z)); o.z = kk.z * (dz - m1.z - yh * m2.z);
                yh = (yv[k].w - mu.w) * is.w; dz = gv[k].w * gelu_grad_f(fmaf(yh, ga.w, be.w)); o.w = kk.w * (dz - m1.w - yh * m2.w);
                *reinterpret_cast<float4*>(dy + rr * C + c) = o;
                am = amax4(am, o);
            }
        }
    }
    if (amax) block_amax_commit(am, amax);
}

// GLU backward, column-stationary (H % 4 == 0, H/4 <= 256): thread (cx, ry) owns a-columns 4cx.. and the matching gate columns;
// optionally accumulates the bias gradient of the GLU convolution, dbias[c] = sum_rows dh[row][c] (2H columns), from the values
// it has just produced (shared-memory combine over the row lanes, one fp32 atomic per column per block; zeroed by the caller).
__global__ void __launch_bounds__(320, 2)
glu_bwd_cs_kernel(const float* __restrict__ g, const float* __restrict__ h, float* __restrict__ dh, float* __restrict__ dbias,
                  long long rows, int H, unsigned int* __restrict__ amax) {
    extern __shared__ float red[];                       // [RY][2H]
    const int c = threadIdx.x * 4, RY = blockDim.y;
    const long long r0 = (long long)blockIdx.x * CS_ROWS_PER_BLOCK;
    const long long r1 = min(rows, r0 + CS_ROWS_PER_BLOCK);
    float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
    float am = 0.f;
    for (long long r = r0 + threadIdx.y; r < r1; r += (long long)CS_U * RY) {
        float4 av[CS_U], bv[CS_U], gv[CS_U];
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) { av[k] = ld4s(h + rr * 2 * H + c); bv[k] = ld4s(h + rr * 2 * H + H + c); gv[k] = ld4s(g + rr * H + c); }
        }
#pragma unroll
        for (int k = 0; k < CS_U; ++k) {
            const long long rr = r + (long long)k * RY;
            if (rr < r1) {
                const float4 a = av[k], b = bv[k], gg = gv[k];
                float4 da, db;
                float s;
                s = sigmoid_fast(b.x); da.x = gg.x * s; db.x = gg.x * a.x * s * (1.f - s);
                s = sigmoid_fast(b.y); da.y = gg.y * s; db.y = gg.y * a.y * s * (1.f - s);
                s = sigmoid_fast(b.z); da.z = gg.z * s; db.z = gg.z * a.z * s * (1.f - s);
                s = sigmoid_fast(b.w); da.w = gg.w * s; db.w = gg.w * a.w * s * (1.f - s);
                *reinterpret_cast<float4*>(dh + rr * 2 * H + c) = da;
                *reinterpret_cast<float4*>(dh + rr * 2 * H + H + c) = db;
                sa.x += da.x; sa.y += da.y; sa.z += da.z; sa.w += da.w;
                sb.x += db.x; sb.y += db.y; sb.z += db.z; sb.w += db.w;
                am = amax4(amax4(am, da), db);
            }
        }
    }
    if (amax) block_amax_commit(am, amax);
    if (!dbias) return;
    float* mine = red + (size_t)threadIdx.y * 2 * H;
    *reinterpret_cast<float4*>(mine + c) = sa;
    *reinterpret_cast<float4*>(mine + H + c) = sb;
    __syncthreads();
    const int tid = threadIdx.y * blockDim.x + threadIdx.x, nthr = blockDim.x * blockDim.y;
    for (int i = tid; i < 2 * H; i += nthr) {
        float acc = 0.f;
        for (int ry = 0; ry < RY; ++ry) acc += red[(size_t)ry * 2 * H + i];
        atomicAdd(dbias + i, acc);
    }
}

inline bool aligned16(const void* a, const void* b = nullptr, const void* c = nullptr) {
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c)) & 15) == 0;
}

inline int ew_grid(long long total, int block = 256, int per_thread = 1) {
    long long g = (total / per_thread + block - 1) / block;
    long long cap = (long long)num_sms() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace bm
