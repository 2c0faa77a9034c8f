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
// Two kernels forward (column statistics, then one THREAD per pixel over TMA-staged tiles), one backward:
//   d total / d out[b,(y,x),c] = cv * valid * (softmax_c - onehot) + ct * (cnt[b,x,c] * softmax_y - onehot)
// with cnt[b,x,c] = #{y : label[b,y,x] == c}; cnt * softmax_y = exp(out - adj), adj = lse_y - ln(cnt) (+inf where
// cnt = 0), which is what the column kernel stores next to lse_y.  fp32, max-subtracted like ATen's log_softmax; exp / log are
// ex2.approx(fma(v, log2e, -m*log2e)) / lg2.approx (the accurate expf costs ~18 instructions and made all three
// kernels issue-bound: ncu, profiles/r01_k4_contrastive.md); the loss sums are accumulated in double.
#include "common.cuh"
#include "tc_common.cuh"

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
              float* __restrict__ col_lse, float* __restrict__ col_adj) {
    extern __shared__ __align__(16) int s_lab[]; // the column's labels as int (-1 = not a class id), h + K4_CU slots
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
    const int stride = w * C;                    // B*h*w*C < 2^31 (checked by the host entry)
    float m = -INFINITY, s = 0.f;
    int cnt = 0;
    for (int y0 = 0; y0 < h; y0 += K4_CU, p += K4_CU * stride) {
        float v[K4_CU];
        if (y0 + K4_CU <= h) {
#pragma unroll
            for (int i = 0; i < K4_CU; ++i) v[i] = p[i * stride];
        } else {
#pragma unroll
            for (int i = 0; i < K4_CU; ++i) v[i] = y0 + i < h ? p[i * stride] : -INFINITY;
        }
        float cm = v[0];
#pragma unroll
        for (int i = 1; i < K4_CU; ++i) cm = fmaxf(cm, v[i]);
        const float mn = fmaxf(m, cm), nmn = -mn * K4_L2E;
        float cs = 0.f;
#pragma unroll
        for (int i = 0; i < K4_CU; i += 4) {
            const int4 l4 = *reinterpret_cast<const int4*>(&s_lab[y0 + i]);
            cs += k4_exp(v[i], nmn) + k4_exp(v[i + 1], nmn);             // exp(-inf) = 0 for the padding rows
            cs += k4_exp(v[i + 2], nmn) + k4_exp(v[i + 3], nmn);
            cnt += (l4.x == c) + (l4.y == c) + (l4.z == c) + (l4.w == c);
        }
        s = s * k4_exp(m, nmn) + cs;             // first chunk: s = 0, exp(-inf) = 0
        m = mn;
    }
    const size_t o = ((size_t)b * w + x) * C + c;
    const float lse = fmaf(lg2f(s), K4_LN2, m);
    col_lse[o] = lse;
    col_adj[o] = cnt ? fmaf(-lg2f((float)cnt), K4_LN2, lse) : INFINITY;
}

// ---- one thread per pixel over shared-memory tiles ------------------------------------------------------------
// A tile is TP consecutive pixels = TP*C contiguous floats (TP % 4 == 0 keeps every tile 16-byte aligned and sized), so
// one 1-D bulk copy (TMA, mbarrier-completed) stages it with no per-element instructions; the backward stages the
// matching [TP, C] slice of adj the same way and bulk-stores the gradient tile it builds in place.  One thread (forward)
// or four (backward: classes r, r+4, ..) walk a row of the tile: no predicates, no per-element addressing; a warp-per-pixel version
// of this kernel was issue-bound at 8x the instructions (profiles/r01_k4_contrastive.md).  Several CTAs per SM overlap
// the copy of one tile with the arithmetic of another.  Ragged cases (last partial tile, w % TP != 0 for the backward,
// unaligned pointers) take a cooperative-copy path.
__device__ __forceinline__ void k4_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(tc::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void k4_bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(gdst), "r"(tc::smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void k4_bulk_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void k4_fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// R threads per pixel (classes r, r + R, ...): the backward, with two staged tiles per pixel, needs R = 4 to have
// enough warps per SM; the forward is fastest with one thread per pixel (measured: 21 vs 34 us at B=8, h=128).
template <bool BWD, int TP, int R>
__global__ void __launch_bounds__(TP * R)
k4_px_kernel(const float* __restrict__ out, const long long* __restrict__ labels, int n_px, int hw, int w,
             int C, long long ignore, const float* __restrict__ col_lse, const float* __restrict__ col_adj,
             double* __restrict__ loss_sums, unsigned long long* __restrict__ counts,
             const float* __restrict__ coef, float* __restrict__ grad, int fast) {
    constexpr int NT = TP * R;
    extern __shared__ __align__(128) float k4_smem[];
    __shared__ __align__(8) uint64_t bar;
    float* s_out = k4_smem;
    float* s_adj = k4_smem + TP * C;
    const int tid = threadIdx.x, pi = tid / R, r = tid % R;
    const int n_tiles = (n_px + TP - 1) / TP;
    const uint32_t tile_bytes = (uint32_t)(TP * C) * 4u;
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_barrier_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    double lv = 0.0, lt = 0.0;
    int nv = 0, bad = 0;
    float cv = 0.f, ct = 0.f;
    if (BWD) {
        cv = coef[0];
        ct = coef[1];
    }
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int px0 = tile * TP;
        const int n = min(TP, n_px - px0);
        const bool tma = fast && n == TP;
        const int px = min(px0 + pi, n_px - 1);
        const long long lab = labels[px];                        // in flight under the tile copy
        const int cb = ((px / hw) * w + px % w) * C;             // hw is a multiple of w
        if (tma) {
            if (tid == 0) {
                tc::mbar_arrive_expect_tx(&bar, BWD ? 2 * tile_bytes : tile_bytes);
                k4_bulk_load(s_out, out + (size_t)px0 * C, tile_bytes, &bar);
                if (BWD) k4_bulk_load(s_adj, col_adj + cb, tile_bytes, &bar);      // thread 0: cb of the tile's first pixel
            }
            tc::mbar_wait(&bar, parity);
            parity ^= 1;
        } else {
            for (int i = tid; i < n * C; i += NT) {
                s_out[i] = out[(size_t)px0 * C + i];
                if (BWD) {
                    const int j = i / C, q = px0 + j;
                    s_adj[i] = col_adj[((q / hw) * w + q % w) * C + (i - j * C)];
                }
            }
            __syncthreads();
        }
        // every thread runs the reductions (rows past the tile's end read a clamped row); only pi < n has effects
        const bool live = pi < n;
        float* row = s_out + min(pi, n - 1) * C;
        float m0 = -INFINITY, m1 = -INFINITY;
        int c = r;
#pragma unroll 4
        for (; c + R < C; c += 2 * R) {
            m0 = fmaxf(m0, row[c]);
            m1 = fmaxf(m1, row[c + R]);
        }
        if (c < C) m0 = fmaxf(m0, row[c]);
        float m = fmaxf(m0, m1);
#pragma unroll
        for (int o = 1; o < R; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        const float nm = -m * K4_L2E;
        float s0 = 0.f, s1 = 0.f;
        c = r;
#pragma unroll 4
        for (; c + R < C; c += 2 * R) {
            s0 += k4_exp(row[c], nm);
            s1 += k4_exp(row[c + R], nm);
        }
        if (c < C) s0 += k4_exp(row[c], nm);
        float sum = s0 + s1;
#pragma unroll
        for (int o = 1; o < R; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float lse = fmaf(lg2f(sum), K4_LN2, m);
        const bool in_range = lab >= 0 && lab < C;
        const bool counted = lab != ignore;
        if (!BWD) {
            if (live && r == 0) {
                if (in_range) {
                    const float ol = row[(int)lab];
                    lt += (double)(col_lse[cb + (int)lab] - ol);
                    if (counted) {
                        lv += (double)(lse - ol);
                        nv += 1;
                    }
                } else {
                    bad += 1;     // F.one_hot raises on such a label (and so does the class-index CE unless ignored)
                }
            }
        } else if (live) {
            const float* arow = s_adj + pi * C;
            const float nl = -lse * K4_L2E;
            const float cvv = counted ? cv : 0.f;
            const int lc = in_range ? (int)lab : -1;
#pragma unroll 4
            for (c = r; c < C; c += R) {
                const float v = row[c];
                const float gt = k4_exp(v, -arow[c] * K4_L2E);          // adj = +inf -> 0
                const float gv = k4_exp(v, nl);
                float g = fmaf(cvv, gv, ct * gt);
                if (c == lc) g -= cvv + ct;
                row[c] = g;
            }
        }
        if (BWD) {
            if (tma) {
                k4_fence_proxy_async();                          // the tile was written through the generic proxy
                __syncthreads();
                if (tid == 0) {
                    k4_bulk_store(grad + (size_t)px0 * C, s_out, tile_bytes);
                    k4_bulk_store_wait_read();                   // smem may be overwritten by the next tile's copy
                }
            } else {
                __syncthreads();
                for (int i = tid; i < n * C; i += NT) grad[(size_t)px0 * C + i] = s_out[i];
            }
        }
        __syncthreads();
    }
    if (!BWD) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lv += __shfl_xor_sync(0xffffffffu, lv, o);
            lt += __shfl_xor_sync(0xffffffffu, lt, o);
            nv += __shfl_xor_sync(0xffffffffu, nv, o);
            bad += __shfl_xor_sync(0xffffffffu, bad, o);
        }
        if ((tid & 31) == 0) {
            atomicAdd(&loss_sums[0], lv);
            atomicAdd(&loss_sums[1], lt);
            if (nv) atomicAdd(&counts[0], (unsigned long long)nv);
            if (bad) atomicAdd(&counts[1], (unsigned long long)bad);
        }
    }
}

constexpr int K4_TP_FWD = 64, K4_R_FWD = 1, K4_TP_BWD = 32, K4_R_BWD = 4;

static int k4_check(const void* a, const void* b, int B, int h, int w, int C) {
    if (B < 0 || h <= 0 || w <= 0 || C <= 0) return fail(LC2IS_ERR_SHAPE, "bad shape%s");
    if (C > K4_CMAX) return fail(LC2IS_ERR_UNSUPPORTED, "contrastive loss kernels hold C <= 160 classes%s");
    if (B > 65535 || (long long)B * h * w * C > 0x7fffffffLL) return fail(LC2IS_ERR_UNSUPPORTED, "batch too large%s");
    if (h > 8192) return fail(LC2IS_ERR_UNSUPPORTED, "h > 8192%s");
    if (B && (!a || !b)) return fail(LC2IS_ERR_ARG, "null pointer%s");
    return 0;
}

template <bool BWD, int TP, int R>
static int k4_launch_px(const float* out, const long long* labels, int n_px, int hw, int w, int C, long long ignore,
                        const float* col_lse, const float* col_adj, double* loss_sums, unsigned long long* counts,
                        const float* coef, float* grad, cudaStream_t st) {
    auto kernel = k4_px_kernel<BWD, TP, R>;
    const size_t smem = (size_t)(BWD ? 2 : 1) * TP * C * sizeof(float);
    static bool attr_done = false;               // same value every time; a race only repeats the call
    if (!attr_done) {
        LC2IS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * TP * K4_CMAX * 4));
        attr_done = true;
    }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TP * R, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    const int n_tiles = (n_px + TP - 1) / TP, cap = sm_count() * per_sm;
    // bulk copies need 16-byte aligned tiles; the backward also needs a tile's pixels inside one image row
    int fast = ((uintptr_t)out % 16 == 0) && (!BWD || ((uintptr_t)col_adj % 16 == 0 && (uintptr_t)grad % 16 == 0 &&
                                                        w % TP == 0));
    kernel<<<n_tiles < cap ? n_tiles : cap, TP * R, smem, st>>>(out, labels, n_px, hw, w, C, ignore, col_lse, col_adj,
                                                             loss_sums, counts, coef, grad, fast);
    return 0;
}

}  // namespace lc2is

using namespace lc2is;

extern "C" int lc2is_contrastive_fwd(const float* d_out, const int64_t* d_labels, int B, int h, int w, int C,
                                     int64_t ignore_index, float* d_col_lse, float* d_col_adj,
                                     double* d_loss_sums, int64_t* d_counts, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = k4_check(d_out, d_labels, B, h, w, C)) return e;
    if (B == 0) return 0;
    if (!d_col_lse || !d_col_adj || !d_loss_sums || !d_counts) return fail(LC2IS_ERR_ARG, "null pointer%s");
    k4_col_kernel<<<dim3(w, B), K4_CMAX, (size_t)(h + K4_CU) * sizeof(int), (cudaStream_t)stream>>>(
        d_out, (const long long*)d_labels, h, w, C, d_col_lse, d_col_adj);
    LC2IS_CHECK_LAUNCH("k4_col_kernel");
    if (int e = k4_launch_px<false, K4_TP_FWD, K4_R_FWD>(d_out, (const long long*)d_labels, B * h * w, h * w, w, C,
                                               (long long)ignore_index, d_col_lse, d_col_adj, d_loss_sums,
                                               (unsigned long long*)d_counts, nullptr, nullptr, (cudaStream_t)stream))
        return e;
    LC2IS_CHECK_LAUNCH("k4_px_kernel<fwd>");
    return 0;
}

extern "C" int lc2is_contrastive_bwd(const float* d_out, const int64_t* d_labels, int B, int h, int w, int C,
                                     int64_t ignore_index, const float* d_col_lse, const float* d_col_adj,
                                     const float* d_coef, float* d_grad, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = k4_check(d_out, d_labels, B, h, w, C)) return e;
    if (B == 0) return 0;
    if (!d_col_lse || !d_col_adj || !d_coef || !d_grad) return fail(LC2IS_ERR_ARG, "null pointer%s");
    if (int e = k4_launch_px<true, K4_TP_BWD, K4_R_BWD>(d_out, (const long long*)d_labels, B * h * w, h * w, w, C,
                                              (long long)ignore_index, d_col_lse, d_col_adj, nullptr, nullptr, d_coef,
                                              d_grad, (cudaStream_t)stream))
        return e;
    LC2IS_CHECK_LAUNCH("k4_px_kernel<bwd>");
    return 0;
}
