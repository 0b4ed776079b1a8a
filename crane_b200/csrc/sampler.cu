// Device-side sampler: what crane-serve runs on the logits of every decode step
// (`sampling::sample`, crane-serve/src/engine/sampling.rs:169-380), here fused behind the lm_head so that only a token id
// leaves the GPU:
//   penalties   `apply_penalties_inplace` (sampling.rs:422-480): per DISTINCT token of the trailing context window,
//               logit >= 0 ? logit / rp : logit * rp, then minus (count * frequency + presence);
//   top-k       `topk_indices` (crane-core/src/ops/fused_ops/portable.rs:28-66, kernels/cuda/topk.cu:213-259): the k largest in the
//               TOTAL order (value descending, index ascending) -- radix select of the k-th value, ties resolved by ascending
//               index, then a bitonic sort of the k (value, index) keys;
//   top-p       (sampling.rs:312-330) softmax(topk / T), running sum, keep j while cumsum_j <= p or cumsum_(j-1) <= p (the first
//               token that crosses p is kept), the others masked to -1e9;
//   draw        `sample_gumbel_max_idx` (sampling.rs:382-393): argmax(l / T - log(-log u)), u ~ U(1e-7, 0.999).  The
//               reference draws u with candle's device RNG, whose stream no test pins: the uniforms are an INPUT here (the
//               caller passes them, or a seed for the built-in counter hash).
// HBM-bound integer/compare work: one CTA per logits row; a 1 MB row is re-read from L2 by the three 11/11/10-bit select passes.
#include "sampler.cuh"

namespace cb {

__device__ __forceinline__ uint32_t f2ord(float f) {           // monotone: larger float <=> larger unsigned
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ unsigned long long topk_key(uint32_t ord, uint32_t idx) {     // ascending key = value descending, index ascending
    return ((unsigned long long)(~ord) << 32) | (unsigned long long)idx;
}
// U(1e-7, 0.999) from a counter hash (splitmix64) -- stands in for candle's `rand_like` when the caller passes no uniforms
__device__ __forceinline__ float hash_uniform(unsigned long long seed, unsigned long long i) {
    unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (i + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float r = ((float)(z >> 40) + 0.5f) * (1.0f / 16777216.0f);
    return 1e-7f + (0.999f - 1e-7f) * r;
}

// ------------------------------------------------------------------------------------------------ penalties
__global__ void __launch_bounds__(256)
sampler_penalties_kernel(float* __restrict__ logits, int V, const SamplerRow* __restrict__ rows, const uint32_t* __restrict__ ctx) {
    const SamplerRow r = rows[blockIdx.x];
    const bool rep = r.rep_mul_pos != 1.f || r.rep_mul_neg != 1.f;
    const bool fp = r.frequency_penalty != 0.f || r.presence_penalty != 0.f;
    if (r.ctx_len == 0 || (!rep && !fp)) return;
    const uint32_t* c = ctx + r.ctx_off;
    float* row = logits + (size_t)blockIdx.x * V;
    for (int i = threadIdx.x; i < r.ctx_len; i += blockDim.x) {
        const uint32_t t = c[i];
        int cnt = 0;
        bool first = true;
        for (int j = 0; j < r.ctx_len; ++j)
            if (c[j] == t) { ++cnt; if (j < i) first = false; }
        if (!first || t >= (uint32_t)V) continue;            // one update per distinct token
        float l = row[t];
        if (rep) l = (l >= 0.f) ? l * r.rep_mul_pos : l * r.rep_mul_neg;
        if (fp) l -= (float)cnt * r.frequency_penalty + r.presence_penalty;
        row[t] = l;
    }
}

int sampler_penalties_launch(cudaStream_t st, float* logits, int V, int rows, const SamplerRow* rows_dev, const uint32_t* ctx_dev) {
    return launch_k(sampler_penalties_kernel, dim3(rows), dim3(256), 0, st, false, logits, V, rows_dev, ctx_dev);
}

// ------------------------------------------------------------------------------------------------ top-k
constexpr int TK_THREADS = 1024;

__global__ void __launch_bounds__(TK_THREADS)
sampler_topk_kernel(const float* __restrict__ logits, int V, int k, uint32_t* __restrict__ idx_out, float* __restrict__ val_out) {
    __shared__ unsigned int hist[2048];
    __shared__ unsigned long long keys[TOPK_MAX];
    __shared__ uint32_t s_prefix, s_mask, s_remaining, s_ngt, s_base;
    __shared__ uint32_t wc[TK_THREADS / 32];
    const float* x = logits + (size_t)blockIdx.x * V;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_prefix = 0; s_mask = 0; s_remaining = (uint32_t)k; }
    __syncthreads();
    // ---- radix select of the k-th largest value: 11 + 11 + 10 bits, most significant first ----
    for (int pass = 0; pass < 3; ++pass) {
        const int shift = pass == 0 ? 21 : pass == 1 ? 10 : 0;
        const int nb = pass == 2 ? 1024 : 2048;
        for (int i = tid; i < nb; i += TK_THREADS) hist[i] = 0;
        __syncthreads();
        const uint32_t prefix = s_prefix, mask = s_mask, remaining = s_remaining;
        for (int i = tid; i < V; i += TK_THREADS) {
            const uint32_t u = f2ord(x[i]);
            if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & (uint32_t)(nb - 1)], 1u);
        }
        __syncthreads();
        if (warp == 0) {             // lane L owns the bins [nb - (L+1) per, nb - L per): lane 0 holds the largest digits
            const int per = nb / 32;
            const int hi = nb - lane * per;
            uint32_t lsum = 0;
            for (int b = hi - per; b < hi; ++b) lsum += hist[b];
            uint32_t incl = lsum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            const uint32_t excl = incl - lsum;
            if (excl < remaining && incl >= remaining) {
                uint32_t c = excl;
                for (int b = hi - 1; b >= hi - per; --b) {
                    const uint32_t c2 = c + hist[b];
                    if (c2 >= remaining) {
                        s_prefix = prefix | ((uint32_t)b << shift);
                        s_mask = mask | ((uint32_t)(nb - 1) << shift);
                        s_remaining = remaining - c;         // how many of this digit are still needed
                        break;
                    }
                    c = c2;
                }
            }
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;                 // bit pattern of the k-th largest value
    const uint32_t need_eq = s_remaining;          // elements equal to it that belong to the top k: the ones with the LOWEST indices
    const uint32_t n_gt = (uint32_t)k - need_eq;   // strictly larger elements: all of them
    if (tid == 0) { s_ngt = 0; s_base = 0; }
    __syncthreads();
    for (int c0 = 0; c0 < V; c0 += TK_THREADS) {
        const int i = c0 + tid;
        const uint32_t u = i < V ? f2ord(x[i]) : 0u;
        const bool gt = i < V && u > thr;
        const bool eq = i < V && u == thr;
        if (gt) keys[atomicAdd(&s_ngt, 1u)] = topk_key(u, (uint32_t)i);
        const uint32_t base = s_base;
        if (base < need_eq) {                      // CTA-uniform: rank the equals of this chunk in index order
            const uint32_t bal = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) wc[warp] = __popc(bal);
            __syncthreads();
            uint32_t woff = 0, total = 0;
            for (int w = 0; w < TK_THREADS / 32; ++w) { const uint32_t c = wc[w]; if (w < warp) woff += c; total += c; }
            const uint32_t rank = base + woff + __popc(bal & ((1u << lane) - 1u));
            if (eq && rank < need_eq) keys[n_gt + rank] = topk_key(u, (uint32_t)i);
            __syncthreads();
            if (tid == 0) s_base = base + total;
            __syncthreads();
        }
    }
    // ---- sort the k keys (padded to a power of two with +inf keys) ----
    int P = 1;
    while (P < k) P <<= 1;
    __syncthreads();
    for (int i = k + tid; i < P; i += TK_THREADS) keys[i] = ~0ull;
    for (int kk = 2; kk <= P; kk <<= 1)
        for (int j = kk >> 1; j > 0; j >>= 1) {
            __syncthreads();
            for (int i = tid; i < P; i += TK_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], b = keys[ixj];
                    const bool asc = (i & kk) == 0;
                    if ((a > b) == asc) { keys[i] = b; keys[ixj] = a; }
                }
            }
        }
    __syncthreads();
    for (int i = tid; i < k; i += TK_THREADS) {
        const uint32_t idx = (uint32_t)(keys[i] & 0xffffffffull);
        idx_out[(size_t)blockIdx.x * k + i] = idx;
        if (val_out) val_out[(size_t)blockIdx.x * k + i] = x[idx];
    }
}

int sampler_topk_launch(cudaStream_t st, const float* logits, int V, int rows, int k, uint32_t* idx_out, float* val_out) {
    if (k < 1 || k > TOPK_MAX || k > V || rows < 1) return -1000;
    return launch_k(sampler_topk_kernel, dim3(rows), dim3(TK_THREADS), 0, st, false, logits, V, k, idx_out, val_out);
}

// ------------------------------------------------------------------------------------------------ draw
__global__ void __launch_bounds__(256)
sampler_draw_kernel(const float* __restrict__ logits, int V, const SamplerRow* __restrict__ rows, const uint32_t* __restrict__ topk_idx,
                    const float* __restrict__ topk_val, int k_stride, const float* __restrict__ uniforms, uint32_t* __restrict__ tokens_out) {
    __shared__ float sv[32];
    __shared__ int si[32];
    const SamplerRow r = rows[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t* idx = topk_idx + (size_t)blockIdx.x * k_stride;
    const float* val = topk_val + (size_t)blockIdx.x * k_stride;
    if (r.temperature <= 0.f) {                     // greedy: the head of the top-k order == lowest index among the maxima
        if (tid == 0) tokens_out[blockIdx.x] = idx[0];
        return;
    }
    // `logits / temperature` is candle's affine op: a multiplication by the f64 reciprocal rounded to f32
    const float inv_t = (float)(1.0 / (double)r.temperature);
    const bool scale = r.temperature != 1.f;
    auto uni = [&](int j) { return r.uni_off >= 0 ? uniforms[r.uni_off + j] : hash_uniform(r.seed, (unsigned long long)j); };
    if (r.top_k > 0) {
        if (tid != 0) return;                       // <= 64 candidates: one thread, in the reference's operation order
        const int k = r.top_k;
        float masked[64];
        const bool top_p_on = r.top_p > 0.f && r.top_p < 1.f;
        if (top_p_on) {
            float m = -INFINITY;
            for (int j = 0; j < k; ++j) m = fmaxf(m, val[j] * inv_t);
            float sum = 0.f;
            for (int j = 0; j < k; ++j) sum += expf(val[j] * inv_t - m);
            float cum = 0.f;
            bool prev_le = false;
            for (int j = 0; j < k; ++j) {
                cum += expf(val[j] * inv_t - m) / sum;
                const bool le = cum <= r.top_p;
                masked[j] = (le || prev_le) ? val[j] : -1e9f;
                prev_le = le;
            }
        } else {
            for (int j = 0; j < k; ++j) masked[j] = val[j];
        }
        float best = -INFINITY;
        int pos = 0;
        for (int j = 0; j < k; ++j) {
            const float sc = (scale ? masked[j] * inv_t : masked[j]) - logf(-logf(uni(j)));
            if (sc > best) { best = sc; pos = j; }
        }
        tokens_out[blockIdx.x] = idx[pos];
        return;
    }
    // whole vocabulary (no top-k, no top-p): argmax_i (l_i / T - log(-log u_i)), first maximum
    const float* x = logits + (size_t)blockIdx.x * V;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += 256) {
        const float sc = (scale ? x[i] * inv_t : x[i]) - logf(-logf(uni(i)));
        if (sc > bv) { bv = sc; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if (lane == 0) { sv[warp] = bv; si[warp] = bi; }
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < 8; ++w)
            if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
        tokens_out[blockIdx.x] = (unsigned)bi < (unsigned)V ? (uint32_t)bi : 0u;
    }
}

int sampler_draw_launch(cudaStream_t st, const float* logits, int V, int rows, const SamplerRow* rows_dev, const uint32_t* topk_idx,
                        const float* topk_val, int k_stride, const float* uniforms_dev, uint32_t* tokens_out) {
    return launch_k(sampler_draw_kernel, dim3(rows), dim3(256), 0, st, false, logits, V, rows_dev, topk_idx, topk_val, k_stride, uniforms_dev,
                    tokens_out);
}

}  // namespace cb
