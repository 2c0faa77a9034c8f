// K4: ContrastiveLoss (reference model/loss.py:39-64) on the low-resolution logits [B, h*w, C].
//
//   loss_visual  = CrossEntropy over the class axis, per pixel, ignore_index honoured (loss.py:52,55,59):
//                    mean over counted pixels of  lse_c(out[b,p,:]) - out[b,p,label]
//   loss_textual = CrossEntropy of the [B,h,w,C] view against the one-hot [B,h,w,151] float target (loss.py:51,54,58).
//                  torch takes dim 1 as the class axis, i.e. the softmax runs over the image-ROW axis y for every
//                  (b, x, c) column, and 'mean' divides by B*w*C:
//                    1/(B w C) * sum_{b,y,x} ( lse_y(out[b,(:,x),label]) - out[b,(y,x),label] )
//   total        = (loss_textual + loss_visual) / 2                                              (loss.py:64)
//
// Two kernels forward (column statistics, then one warp per pixel), one backward (one warp per pixel):
//   d total / d out[b,(y,x),c] = cv * valid * (softmax_c - onehot) + ct * (cnt[b,x,c] * softmax_y - onehot)
// with cnt[b,x,c] = #{y : label[b,y,x] == c}.  fp32, max-subtracted like ATen's log_softmax; exp / log are
// ex2.approx(fma(v, log2e, -m*log2e)) / lg2.approx (the accurate expf costs ~18 instructions and made all three
// kernels issue-bound: ncu, profiles/r01_k4_contrastive.md); the loss sums are accumulated in double.
#include "common.cuh"

namespace lc2is {

constexpr int K4_NR = 5;                 // classes per lane: C <= 160
constexpr int K4_CMAX = 32 * K4_NR;
constexpr int K4_WARPS = 8;
constexpr float K4_L2E = 1.4426950408889634f, K4_LN2 = 0.6931471805599453f;
// exp(v - m) given nm = -m * log2(e)
__device__ __forceinline__ float k4_exp(float v, float nm) { return ex2f(fmaf(v, K4_L2E, nm)); }

// ---- column statistics: lse over y and label counts, one CTA per (b, x), one thread per class ------------------
// One pass with K4_CU rows in flight per thread (the loads of a thread are w*C*4 bytes apart, the threads of a warp
// read 128 contiguous bytes): running (max, sum) pair, rescaled once per chunk.
constexpr int K4_CU = 16;
__global__ void __launch_bounds__(K4_CMAX, 8)
k4_col_kernel(const float* __restrict__ out, const long long* __restrict__ labels, int h, int w, int C,
              float* __restrict__ col_lse, float* __restrict__ col_cnt) {
    extern __shared__ int s_lab[];               // the column's labels as int (-1 = not a class id), h + K4_CU slots
    const int x = blockIdx.x, b = blockIdx.y, c = threadIdx.x;
    const long long* lab = labels + (size_t)b * h * w + x;
    const int hp = (h + K4_CU - 1) / K4_CU * K4_CU;
    for (int y = c; y < hp; y += K4_CMAX) {
        long long l = y < h ? lab[(size_t)y * w] : -1;
        s_lab[y] = (l >= 0 && l < C) ? (int)l : -1;
    }
    __syncthreads();
    if (c >= C) return;
    const float* p = out + ((size_t)b * h * w + x) * C + c;
    const size_t stride = (size_t)w * C;
    float m = -INFINITY, s = 0.f;
    int cnt = 0;
    for (int y0 = 0; y0 < h; y0 += K4_CU) {
        float v[K4_CU];
#pragma unroll
        for (int i = 0; i < K4_CU; ++i) v[i] = y0 + i < h ? p[(size_t)(y0 + i) * stride] : -INFINITY;
        float cm = v[0];
#pragma unroll
        for (int i = 1; i < K4_CU; ++i) cm = fmaxf(cm, v[i]);
        const float mn = fmaxf(m, cm), nmn = -mn * K4_L2E;
        float cs = 0.f;
#pragma unroll
        for (int i = 0; i < K4_CU; ++i) {
            cs += k4_exp(v[i], nmn);             // exp(-inf) = 0 for the padding rows
            cnt += (s_lab[y0 + i] == c);
        }
        s = s * k4_exp(m, nmn) + cs;             // first chunk: s = 0, exp(-inf) = 0
        m = mn;
    }
    const size_t o = ((size_t)b * w + x) * C + c;
    col_lse[o] = fmaf(lg2f(s), K4_LN2, m);
    col_cnt[o] = (float)cnt;
}

// ---- one warp per pixel: the class-axis softmax, both loss terms (forward) or the gradient (backward) ---------
// Persistent warps, PPW consecutive pixels per iteration (their shuffle / exp chains interleave) with the next
// iteration's rows (and, backward, their column statistics) already in flight while the current ones are reduced:
// what bounds this kernel is the number of bytes a warp keeps in flight.
template <bool BWD>
struct K4Row {
    float v[K4_NR], cl[K4_NR], cc[K4_NR];
    long long lab;
    int cb;                                      // (b * w + x) * C
};
template <bool BWD>
__device__ __forceinline__ void k4_load(K4Row<BWD>& r, const float* __restrict__ out,
                                        const long long* __restrict__ labels, const float* __restrict__ col_lse,
                                        const float* __restrict__ col_cnt, int px, int cb, int C, int lane) {
    const float* row = out + (size_t)px * C;
    r.cb = cb;
#pragma unroll
    for (int k = 0; k < K4_NR; ++k) {
        const int c = lane + 32 * k;
        r.v[k] = c < C ? row[c] : -INFINITY;
        if (BWD) {
            r.cl[k] = c < C ? col_lse[cb + c] : 0.f;
            r.cc[k] = c < C ? col_cnt[cb + c] : 0.f;
        }
    }
    r.lab = labels[px];
}
// PPW consecutive pixels from px0 (clamped to the last pixel): one division pair, then x walks along the row
template <bool BWD, int PPW>
__device__ __forceinline__ void k4_load_group(K4Row<BWD>* r, const float* __restrict__ out,
                                              const long long* __restrict__ labels, const float* __restrict__ col_lse,
                                              const float* __restrict__ col_cnt, int px0, int n_px, int hw, int w,
                                              int C, int lane) {
    int b = px0 / hw, x = px0 % w;               // hw is a multiple of w
#pragma unroll
    for (int j = 0; j < PPW; ++j) {
        const int px = min(px0 + j, n_px - 1);
        if (px0 + j > n_px - 1) {                // clamped (warp-uniform): same pixel as n_px - 1
            b = (n_px - 1) / hw;
            x = (n_px - 1) % w;
        }
        k4_load<BWD>(r[j], out, labels, col_lse, col_cnt, px, (b * w + x) * C, C, lane);
        if (++x == w) {
            x = 0;
            b = (px0 + j + 1) / hw;
        }
    }
}

template <bool BWD, int PPW>
__global__ void __launch_bounds__(32 * K4_WARPS)
k4_row_kernel(const float* __restrict__ out, const long long* __restrict__ labels, int n_px, int hw, int w,
              int C, long long ignore, const float* __restrict__ col_lse, const float* __restrict__ col_cnt,
              double* __restrict__ loss_sums, unsigned long long* __restrict__ counts,
              const float* __restrict__ coef, float* __restrict__ grad) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int stride = gridDim.x * K4_WARPS * PPW;
    __shared__ double s_lv[K4_WARPS], s_lt[K4_WARPS];
    __shared__ int s_nv[K4_WARPS], s_bad[K4_WARPS];
    double lv = 0.0, lt = 0.0;
    int nv = 0, bad = 0;
    float cv = 0.f, ct = 0.f;
    if (BWD) {
        cv = coef[0];
        ct = coef[1];
    }
    int px0 = (blockIdx.x * K4_WARPS + wid) * PPW;
    K4Row<BWD> cur[PPW], nxt[PPW];
    if (px0 < n_px) {
        k4_load_group<BWD, PPW>(cur, out, labels, col_lse, col_cnt, px0, n_px, hw, w, C, lane);
    }
    for (; px0 < n_px; px0 += stride) {
        const int pn0 = px0 + stride < n_px ? px0 + stride : px0;      // last round re-loads its own rows (L1 hits)
        k4_load_group<BWD, PPW>(nxt, out, labels, col_lse, col_cnt, pn0, n_px, hw, w, C, lane);
        float m[PPW], s[PPW], lse[PPW];
#pragma unroll
        for (int j = 0; j < PPW; ++j) {
            m[j] = cur[j].v[0];
#pragma unroll
            for (int k = 1; k < K4_NR; ++k) m[j] = fmaxf(m[j], cur[j].v[k]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < PPW; ++j) m[j] = fmaxf(m[j], __shfl_xor_sync(0xffffffffu, m[j], o));
        }
#pragma unroll
        for (int j = 0; j < PPW; ++j) {
            s[j] = 0.f;
#pragma unroll
            for (int k = 0; k < K4_NR; ++k) s[j] += k4_exp(cur[j].v[k], -m[j] * K4_L2E);    // padding lanes: -inf -> 0
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int j = 0; j < PPW; ++j) s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
        }
#pragma unroll
        for (int j = 0; j < PPW; ++j) lse[j] = fmaf(lg2f(s[j]), K4_LN2, m[j]);
#pragma unroll
        for (int j = 0; j < PPW; ++j) {
            const int px = px0 + j;
            if (px >= n_px) break;                                                    // warp-uniform
            const long long lab = cur[j].lab;
            const bool in_range = lab >= 0 && lab < C;
            const bool counted = lab != ignore;
            if (!BWD) {
                if (in_range) {
                    // the lane that owns the label's class holds out[label]
                    const int lc = (int)lab;
                    float ol = 0.f;
#pragma unroll
                    for (int k = 0; k < K4_NR; ++k)
                        if (lane + 32 * k == lc) ol = cur[j].v[k];
                    ol = __shfl_sync(0xffffffffu, ol, lc & 31);
                    if (lane == 0) {
                        lt += (double)(col_lse[cur[j].cb + lc] - ol);
                        if (counted) {
                            lv += (double)(lse[j] - ol);
                            nv += 1;
                        }
                    }
                } else if (lane == 0) {
                    bad += 1;     // F.one_hot raises on such a label (and so does the class-index CE unless ignored)
                }
            } else {
                float* grow = grad + (size_t)px * C;
                const float nl = -lse[j] * K4_L2E;
#pragma unroll
                for (int k = 0; k < K4_NR; ++k) {
                    const int c = lane + 32 * k;
                    if (c < C) {
                        const float oh = (in_range && c == (int)lab) ? 1.f : 0.f;
                        float g = ct * (cur[j].cc[k] * k4_exp(cur[j].v[k], -cur[j].cl[k] * K4_L2E) - oh);
                        if (counted) g += cv * (k4_exp(cur[j].v[k], nl) - oh);
                        grow[c] = g;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PPW; ++j) cur[j] = nxt[j];
    }
    if (!BWD) {
        if (lane == 0) {
            s_lv[wid] = lv;
            s_lt[wid] = lt;
            s_nv[wid] = nv;
            s_bad[wid] = bad;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double a = 0.0, t = 0.0;
            int n = 0, bd = 0;
#pragma unroll
            for (int i = 0; i < K4_WARPS; ++i) {
                a += s_lv[i];
                t += s_lt[i];
                n += s_nv[i];
                bd += s_bad[i];
            }
            atomicAdd(&loss_sums[0], a);
            atomicAdd(&loss_sums[1], t);
            if (n) atomicAdd(&counts[0], (unsigned long long)n);
            if (bd) atomicAdd(&counts[1], (unsigned long long)bd);
        }
    }
}

constexpr int K4_PPW_FWD = 4, K4_PPW_BWD = 2;

static int k4_check(const void* a, const void* b, int B, int h, int w, int C) {
    if (B < 0 || h <= 0 || w <= 0 || C <= 0) return fail(LC2IS_ERR_SHAPE, "bad shape%s");
    if (C > K4_CMAX) return fail(LC2IS_ERR_UNSUPPORTED, "contrastive loss kernels hold C <= 160 classes%s");
    if (B > 65535 || (long long)B * h * w * C > 0x7fffffffLL) return fail(LC2IS_ERR_UNSUPPORTED, "batch too large%s");
    if (h > 8192) return fail(LC2IS_ERR_UNSUPPORTED, "h > 8192%s");
    if (B && (!a || !b)) return fail(LC2IS_ERR_ARG, "null pointer%s");
    return 0;
}

template <typename K>
static int k4_row_blocks(K kernel, int n_px, int ppw) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32 * K4_WARPS, 0) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    const int need = (n_px + K4_WARPS * ppw - 1) / (K4_WARPS * ppw), cap = sm_count() * per_sm;
    return need < cap ? need : cap;
}

}  // namespace lc2is

using namespace lc2is;

extern "C" int lc2is_contrastive_fwd(const float* d_out, const int64_t* d_labels, int B, int h, int w, int C,
                                     int64_t ignore_index, float* d_col_lse, float* d_col_cnt,
                                     double* d_loss_sums, int64_t* d_counts, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = k4_check(d_out, d_labels, B, h, w, C)) return e;
    if (B == 0) return 0;
    if (!d_col_lse || !d_col_cnt || !d_loss_sums || !d_counts) return fail(LC2IS_ERR_ARG, "null pointer%s");
    k4_col_kernel<<<dim3(w, B), K4_CMAX, (size_t)(h + K4_CU) * sizeof(int), (cudaStream_t)stream>>>(d_out, (const long long*)d_labels, h, w, C,
                                                                     d_col_lse, d_col_cnt);
    LC2IS_CHECK_LAUNCH("k4_col_kernel");
    const int n_px = B * h * w;
    const int blocks = k4_row_blocks(k4_row_kernel<false, K4_PPW_FWD>, n_px, K4_PPW_FWD);
    k4_row_kernel<false, K4_PPW_FWD><<<blocks, 32 * K4_WARPS, 0, (cudaStream_t)stream>>>(
        d_out, (const long long*)d_labels, n_px, h * w, w, C, (long long)ignore_index, d_col_lse, d_col_cnt,
        d_loss_sums, (unsigned long long*)d_counts, nullptr, nullptr);
    LC2IS_CHECK_LAUNCH("k4_row_kernel<fwd>");
    return 0;
}

extern "C" int lc2is_contrastive_bwd(const float* d_out, const int64_t* d_labels, int B, int h, int w, int C,
                                     int64_t ignore_index, const float* d_col_lse, const float* d_col_cnt,
                                     const float* d_coef, float* d_grad, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = k4_check(d_out, d_labels, B, h, w, C)) return e;
    if (B == 0) return 0;
    if (!d_col_lse || !d_col_cnt || !d_coef || !d_grad) return fail(LC2IS_ERR_ARG, "null pointer%s");
    const int n_px = B * h * w;
    const int blocks = k4_row_blocks(k4_row_kernel<true, K4_PPW_BWD>, n_px, K4_PPW_BWD);
    k4_row_kernel<true, K4_PPW_BWD><<<blocks, 32 * K4_WARPS, 0, (cudaStream_t)stream>>>(
        d_out, (const long long*)d_labels, n_px, h * w, w, C, (long long)ignore_index, d_col_lse, d_col_cnt,
        nullptr, nullptr, d_coef, d_grad);
    LC2IS_CHECK_LAUNCH("k4_row_kernel<bwd>");
    return 0;
}
