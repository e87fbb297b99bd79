// Retrieval evaluation: what follows the score GEMM in scripts/run_eval_probs.py:237-307 and
// bm/wer.py:96-116 -- softmax over the candidates, top-k, label matching, per-word aggregation -- fused so that the
// [queries, candidates] probability matrix never has to leave the GPU (the reference copies it to the host, row
// batch by row batch, and runs topk there).
//
// All kernels are HBM/L2-bound row scans: algorithmic traffic = the score row once (4 B per candidate); the k
// selection passes re-read the row from L2/L1.
#pragma once
#include "common.cuh"

namespace bm {

// Total order used for selection: larger value first, ties -> smaller column first.
__device__ __forceinline__ unsigned long long rank_key(float v, int col) {
    unsigned u = __float_as_uint(v);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned)(0x7fffffff - col);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// block-wide reductions (blockDim.x multiple of 32, <= 1024); every thread gets the result
__device__ __forceinline__ float block_max_f(float v, float* red) {
    v = warp_max(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) r = fmaxf(r, red[i]);
    return r;
}
__device__ __forceinline__ float block_sum_f(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int i = 0; i < (blockDim.x >> 5); ++i) r += red[i];     // fixed order -> deterministic
    return r;
}

// One block per query row b.
//   value(o)  = (own_values && o == own_col) ? own_values[b] : vals[b*ld + o]           o in [0, n_cols)
//   label(o)  = (own_labels && o == own_col) ? own_labels[b] : labels[o]
//   is_prob=0: values are scores, p = softmax over the row (bm/losses.py:97-102); row_max / row_sum returned
//   is_prob=1: values are already probabilities (negative entries = "not a column", skipped)
//   top_idx / top_prob [Bn, k]: the k most probable columns, most probable first (-1 / 0 when fewer exist)
//   hit[b]  = rank (0 = best) of the first selected column whose label equals targets[b], -1 if none of the k
//             (top-k' hit for every k' <= k: 0 <= hit < k')                  run_eval_probs.py:253-259, wer.py:107-111
//   soft[b] = sum of p over columns whose label equals targets[b]                        wer.py:114-115
__global__ void retrieval_topk_kernel(const float* __restrict__ vals, long long ld, int n_cols,
                                      const float* __restrict__ own_values, int own_col, int is_prob, int k,
                                      const long long* __restrict__ labels, const long long* __restrict__ own_labels,
                                      const long long* __restrict__ targets, long long* __restrict__ top_idx,
                                      float* __restrict__ top_prob, int* __restrict__ hit, float* __restrict__ soft,
                                      float* __restrict__ row_max, float* __restrict__ row_sum) {
    __shared__ float red[32];
    __shared__ unsigned long long red64[32];
    const int b = blockIdx.x;
    const float* v = vals + (long long)b * ld;
    const bool has_own = own_values != nullptr;
    const float own_v = has_own ? own_values[b] : 0.f;
    auto value = [&](int o) { return (has_own && o == own_col) ? own_v : v[o]; };

    float mx = 0.f, inv = 1.f;
    if (!is_prob) {
        float m = -INFINITY;
        for (int o = threadIdx.x; o < n_cols; o += blockDim.x) m = fmaxf(m, value(o));
        mx = block_max_f(m, red);
        float s = 0.f;
        for (int o = threadIdx.x; o < n_cols; o += blockDim.x) s += expf(value(o) - mx);
        s = block_sum_f(s, red);
        inv = 1.f / s;
        if (threadIdx.x == 0) {
            if (row_max) row_max[b] = mx;
            if (row_sum) row_sum[b] = s;
        }
    }
    auto prob = [&](float x) { return is_prob ? x : expf(x - mx) * inv; };

    const bool match = (labels != nullptr) && (targets != nullptr);
    const long long target = match ? targets[b] : 0;
    const long long own_l = (own_labels != nullptr) ? own_labels[b] : 0;
    auto label = [&](int o) { return (own_labels != nullptr && o == own_col) ? own_l : labels[o]; };

    if (soft) {
        float s = 0.f;
        if (match)
            for (int o = threadIdx.x; o < n_cols; o += blockDim.x) {
                float x = value(o);
                if (label(o) == target && !(is_prob && x < 0.f)) s += prob(x);
            }
        s = block_sum_f(s, red);
        if (threadIdx.x == 0) soft[b] = s;
    }

    // k selection passes: pass j takes the largest key strictly below the key taken by pass j-1
    unsigned long long bound = ~0ull;
    int first_hit = -1;
    for (int j = 0; j < k; ++j) {
        unsigned long long best = 0ull;
        for (int o = threadIdx.x; o < n_cols; o += blockDim.x) {
            float x = value(o);
            if (is_prob && x < 0.f) continue;
            unsigned long long key = rank_key(x, o);
            if (key < bound && key > best) best = key;
        }
        best = warp_max_u64(best);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red64[threadIdx.x >> 5] = best;
        __syncthreads();
        best = red64[0];
        for (int i = 1; i < (blockDim.x >> 5); ++i) best = red64[i] > best ? red64[i] : best;
        if (best == 0ull) {                      // fewer than k columns
            if (threadIdx.x == 0) {
                if (top_idx) top_idx[(long long)b * k + j] = -1;
                if (top_prob) top_prob[(long long)b * k + j] = 0.f;
            }
            continue;
        }
        bound = best;
        const int col = 0x7fffffff - (int)(unsigned)(best & 0xffffffffull);
        if (threadIdx.x == 0) {
            if (top_idx) top_idx[(long long)b * k + j] = col;
            if (top_prob) top_prob[(long long)b * k + j] = prob(value(col));
            if (match && first_hit < 0 && label(col) == target) first_hit = j;
        }
    }
    if (hit && threadIdx.x == 0) hit[b] = first_hit;
}

// Probabilities summed per distinct label ("vocabulary word", bm/wer.py:101-104), without atomics:
//   vocab[b][w] = sum_{j in [seg[w], seg[w+1])} softmax(row b)[perm[j]]        w in [0, V)
// `perm` lists the shared columns grouped by word in increasing column order, so each sum runs in the order of a
// sequential scatter_add.  Column V is the query's private word slot:
//   own_word[b] <  V : the own candidate's probability is added to that word, vocab[b][V] = -1 ("not a column")
//   own_word[b] == V : vocab[b][V] = the own candidate's probability (its word occurs nowhere else)
// Without `own_scores` column V is -1.
__global__ void vocab_probs_kernel(const float* __restrict__ scores, long long ld, const float* __restrict__ own_scores,
                                   const float* __restrict__ row_max, const float* __restrict__ row_sum,
                                   const int* __restrict__ perm, const int* __restrict__ seg, int V,
                                   const int* __restrict__ own_word, float* __restrict__ vocab) {
    const int b = blockIdx.x;
    const float* s = scores + (long long)b * ld;
    const float mx = row_max[b], inv = 1.f / row_sum[b];
    const float p_own = own_scores ? expf(own_scores[b] - mx) * inv : 0.f;
    const int w_own = own_scores ? own_word[b] : -1;
    float* out = vocab + (long long)b * (V + 1);
    for (int w = threadIdx.x; w <= V; w += blockDim.x) {
        if (w == V) {
            out[w] = (w_own == V) ? p_own : -1.f;
            continue;
        }
        float acc = 0.f;
        for (int j = seg[w]; j < seg[w + 1]; ++j) acc += expf(s[perm[j]] - mx) * inv;
        if (w == w_own) acc += p_own;           // the own candidate sits in the LAST column (wer.py:93-94)
        out[w] = acc;
    }
}

// own[b] = <a_b, c_b> / (1e-8 + ||c_b||)  -- the score of an estimate against its own true output, i.e. what the
// last column of the reference's score row holds after `negatives[-1] = output` (wer.py:93, losses.py:91-94).
__global__ void rowdot_scaled_kernel(const float* __restrict__ A, const float* __restrict__ Cm, long long K,
                                     float* __restrict__ own) {
    __shared__ double red[2][32];
    const float* a = A + (long long)blockIdx.x * K;
    const float* c = Cm + (long long)blockIdx.x * K;
    double dot = 0.0, ss = 0.0;
    float d = 0.f, q = 0.f;
    int cnt = 0;
    for (long long i = threadIdx.x; i < K; i += blockDim.x) {
        float x = a[i], y = c[i];
        d = fmaf(x, y, d);
        q = fmaf(y, y, q);
        if (++cnt == 64) { dot += (double)d; ss += (double)q; d = q = 0.f; cnt = 0; }
    }
    dot += (double)d;
    ss += (double)q;
    dot = warp_sum_d(dot);
    ss = warp_sum_d(ss);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = dot; red[1][threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double D = 0.0, S = 0.0;
        for (int i = 0; i < (blockDim.x >> 5); ++i) { D += red[0][i]; S += red[1][i]; }
        own[blockIdx.x] = (float)D * (1.f / (1e-8f + (float)sqrt(S)));
    }
}

// probs[b][0:n_cols] = softmax(scores[b][0:n_cols]) for score rows of stride ld >= n_cols (candidate axis padded to
// the tensor-core tile); ClipLoss.get_probabilities (bm/losses.py:97-102).  One block per row.
__global__ void softmax_rows_ld_kernel(const float* __restrict__ scores, long long ld, int n_cols,
                                       float* __restrict__ probs) {
    __shared__ float red[32];
    const float* s = scores + (long long)blockIdx.x * ld;
    float* p = probs + (long long)blockIdx.x * n_cols;
    float m = -INFINITY;
    for (int o = threadIdx.x; o < n_cols; o += blockDim.x) m = fmaxf(m, s[o]);
    const float mx = block_max_f(m, red);
    float acc = 0.f;
    for (int o = threadIdx.x; o < n_cols; o += blockDim.x) acc += expf(s[o] - mx);
    const float inv = 1.f / block_sum_f(acc, red);
    for (int o = threadIdx.x; o < n_cols; o += blockDim.x) p[o] = expf(s[o] - mx) * inv;
}

}  // namespace bm
