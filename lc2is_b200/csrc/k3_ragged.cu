// k3_ragged: resize (bicubic / bilinear, ANY output size) + argmax + per-image statistics of a RAGGED batch in ONE launch.
//
// Replaces the per-image Python loops of  compute_gt_mIOU  (reference metrics.py:61-79: F.interpolate(bicubic,
// size=sizes[i]) -> Softmax2d -> JaccardIndex per image, against ground truth at each image's ORIGINAL size) and
// generate_masks  (reference utils.py:15-22: bicubic to the original size, argmax(dim=0)) - SURVEY 8f-2.
//
// The batch is described by a table  desc[N][4] = { element offset of image i in the flat label / prediction buffers,
// H_i, W_i, first tile of image i }  (int64).  A TILE is 64 x 32 output pixels of one image (lc2is_ragged_tiles); a CTA
// of 128 threads owns a contiguous range of tiles, a thread a 4 x 4 pixel BLOCK of the tile.
//
// Arithmetic = ATen's (and k3_low_gen_kernel's): horizontal tap chain  r = t0*wx0; r = fma(t_b, wx_b, r)  per source
// row, then the vertical chain over the rows in increasing order.  The 4 x 4 pixels of a block touch at most NR source
// rows / columns (NR = 5 when the image is enlarged >= 3x, else 6), so the block loads an NR x NR window per class and
// every pixel runs its chains over the WHOLE window with zero weights outside its own taps: a zero-weight prefix leaves
// the accumulator at (+-)0 and fma(x, w, 0) rounds like x*w, a zero-weight suffix leaves it unchanged - bit-identical to
// the 4-tap chains for finite logits, at 11-15 fma per pixel and class instead of 16 loads + 20 fma.  A non-finite tap
// anywhere in the window makes the pixel's value non-finite; such pixels (and blocks whose window does not fit: images
// that are shrunk) are re-evaluated with the exact per-pixel chains, so NaN / inf poison exactly the pixels they poison
// in argmax(softmax(interpolate(x))).
#include "common.cuh"

namespace lc2is {

constexpr int RG_BX = 16, RG_BY = 8;                       // blocks per tile
constexpr int RG_TW = RG_BX * 4, RG_TH = RG_BY * 4;         // 64 x 32 pixels
constexpr int RG_THREADS = RG_BX * RG_BY;

struct RaggedParams {
    const float* low;                // [N,C,h,w]
    const long long* desc;           // [N][4]
    const long long* labels;         // flat int64, image i at desc[i][0] (H_i*W_i elements), or null
    unsigned long long* confmat;     // [C,C] or null
    unsigned long long* per_image;   // [N,3,C] (TP, target count, prediction count) or null
    long long* pred;                 // flat int64 like labels, or null
    int N, C, h, w;
    long long ntiles;
};

// Source positions and weights of output index o (ATen: src = scale*(o+0.5)-0.5): `first` = the unclamped source index
// of the first tap, wgt[] the NT tap weights.
template <int MODE>
__device__ __forceinline__ void rg_taps(float scale, int o, int in_size, int& first, float (&wgt)[MODE == 0 ? 2 : 4]) {
    float s = scale * ((float)o + 0.5f) - 0.5f;
    if (MODE == 0) {
        if (s < 0.f) s = 0.f;
        int k = (int)s;
        k = k > in_size - 1 ? in_size - 1 : k;
        float t = s - (float)k;
        t = t < 0.f ? 0.f : (t > 1.f ? 1.f : t);
        wgt[0] = 1.f - t; wgt[1] = t;
        first = k;
    } else {
        const int k = (int)floorf(s);
        float c4[4];
        cubic_coeffs(s - (float)k, c4);
#pragma unroll
        for (int a = 0; a < 4; ++a) wgt[a] = c4[a];
        first = k - 1;
    }
}

// The plan of one axis of a 4-pixel block: kb = first source index of the window, Wt[r][k] = weight of window position k
// for pixel r (0 outside the pixel's own taps).  false: the four pixels need more than NR positions.
template <int MODE, int NR>
__device__ __forceinline__ bool rg_axis_plan(float scale, int o0, int in_size, int& kb, float (&Wt)[4][NR]) {
    constexpr int NT = MODE == 0 ? 2 : 4;
    bool ok = true;
    kb = 0;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int first;
        float wgt[NT];
        rg_taps<MODE>(scale, o0 + r, in_size, first, wgt);
        if (r == 0) kb = first;
        const int d = first - kb;                            // >= 0, non-decreasing in r
        ok &= d + NT <= NR;
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            float v = 0.f;
#pragma unroll
            for (int a = 0; a < NT; ++a)
                if (k == d + a) v = wgt[a];
            Wt[r][k] = v;
        }
    }
    return ok;
}

// exact per-pixel evaluation (k3_low_gen_kernel's chains): the prediction of pixel (y, x)
template <int MODE>
__device__ __noinline__ int rg_pixel_exact(const float* __restrict__ base, int C, int h, int w, float scale_y,
                                           float scale_x, int y, int x) {
    constexpr int NT = MODE == 0 ? 2 : 4;
    int fy, fx;
    float wy[NT], wx[NT];
    rg_taps<MODE>(scale_y, y, h, fy, wy);
    rg_taps<MODE>(scale_x, x, w, fx, wx);
    int iy[NT], ix[NT];
#pragma unroll
    for (int a = 0; a < NT; ++a) {
        iy[a] = clampi(fy + a, 0, h - 1);
        ix[a] = clampi(fx + a, 0, w - 1);
    }
    float best = -INFINITY;
    int idx = 0;
    bool bad = false;
    for (int c = 0; c < C; ++c) {
        const float* pc = base + (size_t)c * h * w;
        float acc = 0.f;
#pragma unroll
        for (int a = 0; a < NT; ++a) {
            float r = __ldg(pc + iy[a] * w + ix[0]) * wx[0];
#pragma unroll
            for (int b = 1; b < NT; ++b) r = fmaf(__ldg(pc + iy[a] * w + ix[b]), wx[b], r);
            acc = a == 0 ? r * wy[0] : fmaf(r, wy[a], acc);
        }
        bad |= !(acc < INFINITY);
        if (acc > best) { best = acc; idx = c; }
    }
    return bad ? 0 : idx;
}

// The 16 predictions of one 4 x 4 block through the NR x NR window.  Returns the mask of pixels that must be
// re-evaluated exactly (non-finite value seen), or 0xffff when the window does not fit.
template <int MODE, int NR>
__device__ __forceinline__ unsigned rg_block(const float* __restrict__ base, int C, int h, int w, float scale_y,
                                             float scale_x, int y0, int x0, int (&idx)[16]) {
    float Wy[4][NR], Wx[4][NR];
    int kby, kbx;
    const bool oky = rg_axis_plan<MODE, NR>(scale_y, y0, h, kby, Wy);
    const bool okx = rg_axis_plan<MODE, NR>(scale_x, x0, w, kbx, Wx);
    if (!(oky && okx)) return 0xffffu;
    int roff[NR], ixc[NR];
#pragma unroll
    for (int k = 0; k < NR; ++k) {
        roff[k] = clampi(kby + k, 0, h - 1) * w;
        ixc[k] = clampi(kbx + k, 0, w - 1);
    }
    float best[16];
    unsigned bad = 0;
#pragma unroll
    for (int q = 0; q < 16; ++q) { best[q] = -INFINITY; idx[q] = 0; }
    const size_t plane = (size_t)h * w;
#pragma unroll 1
    for (int c = 0; c < C; ++c) {
        const float* pc = base + (size_t)c * plane;
        float acc[4][4];
#pragma unroll
        for (int k = 0; k < NR; ++k) {
            const float* prow = pc + roff[k];
            float tap[NR];
#pragma unroll
            for (int m = 0; m < NR; ++m) tap[m] = __ldg(prow + ixc[m]);
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                float hx = tap[0] * Wx[px][0];
#pragma unroll
                for (int m = 1; m < NR; ++m) hx = fmaf(tap[m], Wx[px][m], hx);
#pragma unroll
                for (int r = 0; r < 4; ++r) acc[r][px] = k == 0 ? hx * Wy[r][0] : fmaf(hx, Wy[r][k], acc[r][px]);
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int px = 0; px < 4; ++px) {
                const float v = acc[r][px];
                const int q = r * 4 + px;
                if (!(v < INFINITY)) bad |= 1u << q;
                if (v > best[q]) { best[q] = v; idx[q] = c; }
            }
    }
    return bad;
}

template <int MODE>
__global__ void __launch_bounds__(RG_THREADS)
k3_ragged_kernel(const RaggedParams P) {
    extern __shared__ int rg_stat[];                        // [3][C] counts of the CTA's current image
    const int C = P.C, h = P.h, w = P.w;
    const int tid = threadIdx.x, bx = tid % RG_BX, by = tid / RG_BX;
    const long long t0 = P.ntiles * blockIdx.x / gridDim.x, t1 = P.ntiles * (blockIdx.x + 1) / gridDim.x;
    if (P.per_image) {
        for (int i = tid; i < 3 * C; i += RG_THREADS) rg_stat[i] = 0;
        __syncthreads();
    }
    auto flush = [&](int n) {
        __syncthreads();
        for (int i = tid; i < 3 * C; i += RG_THREADS) {
            const int v = rg_stat[i];
            if (v) {
                atomicAdd(&P.per_image[(size_t)n * 3 * C + i], (unsigned long long)v);
                rg_stat[i] = 0;
            }
        }
        __syncthreads();
    };
    // the image of the first tile: last n with desc[n].first_tile <= t0 (binary search; the table is tiny)
    int n = 0;
    {
        int lo = 0, hi = P.N - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.desc[(size_t)mid * 4 + 3] <= t0) lo = mid; else hi = mid - 1;
        }
        n = lo;
    }
    int cur = -1;
    for (long long tile = t0; tile < t1; ++tile) {
        while (n + 1 < P.N && P.desc[(size_t)(n + 1) * 4 + 3] <= tile) ++n;       // (images without tiles are skipped)
        if (n != cur) {
            if (cur >= 0 && P.per_image) flush(cur);
            cur = n;
        }
        const long long off = P.desc[(size_t)n * 4];
        const int H = (int)P.desc[(size_t)n * 4 + 1], W = (int)P.desc[(size_t)n * 4 + 2];
        const long long tl = tile - P.desc[(size_t)n * 4 + 3];
        const int ntx = (W + RG_TW - 1) / RG_TW;
        const int ty = (int)(tl / ntx), tx = (int)(tl - (long long)ty * ntx);
        const int y0 = ty * RG_TH + by * 4, x0 = tx * RG_TW + bx * 4;
        const float scale_y = (float)h / (float)H, scale_x = (float)w / (float)W;   // ATen, size= given (no scale_factor)
        const float* base = P.low + (size_t)n * C * h * w;
        int idx[16];
        unsigned redo = 0;
        const bool live = y0 < H && x0 < W;
        if (live) {
            // images enlarged >= 3x: four consecutive output pixels span at most two source cells -> 5-wide window
            if (scale_y <= (1.f / 3.f) && scale_x <= (1.f / 3.f))
                redo = rg_block<MODE, 5>(base, C, h, w, scale_y, scale_x, y0, x0, idx);
            else
                redo = rg_block<MODE, 6>(base, C, h, w, scale_y, scale_x, y0, x0, idx);
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int y = y0 + (q >> 2), x = x0 + (q & 3);
            const bool in = live && y < H && x < W;
            int pr = idx[q];
            if (in && ((redo >> q) & 1u)) pr = rg_pixel_exact<MODE>(base, C, h, w, scale_y, scale_x, y, x);
            const size_t pix = (size_t)off + (size_t)y * W + x;
            if (in && P.pred) P.pred[pix] = pr;
            bool valid = false;
            int t = 0;
            if (in && P.labels) {
                const long long lab = P.labels[pix];
                valid = lab >= 0 && lab < C;
                t = (int)lab;
            }
            if (P.per_image && valid) {
                if (t == pr) atomicAdd(&rg_stat[t], 1);
                atomicAdd(&rg_stat[C + t], 1);
                atomicAdd(&rg_stat[2 * C + pr], 1);
            }
            if (P.confmat) {                                 // warp-aggregated 64-bit reductions
                const unsigned act = __ballot_sync(0xffffffffu, valid);
                if (valid) {
                    const int key = t * C + pr;
                    const unsigned m = __match_any_sync(act, key);
                    if ((tid & 31) == __ffs(m) - 1) atomicAdd(&P.confmat[key], (unsigned long long)__popc(m));
                }
            }
        }
    }
    if (cur >= 0 && P.per_image) flush(cur);
}

}  // namespace lc2is

using namespace lc2is;

// Tiles of one H x W image (the unit of the descriptor table's `first tile` column).
extern "C" long long lc2is_ragged_tiles(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return (long long)((H + RG_TH - 1) / RG_TH) * ((W + RG_TW - 1) / RG_TW);
}

extern "C" int lc2is_argmax_confmat_ragged(const float* d_low, int N, int C, int h, int w, int mode,
                                           const int64_t* d_desc, long long n_tiles, const int64_t* d_labels,
                                           int64_t* d_confmat, int64_t* d_per_image, int64_t* d_pred,
                                           lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (N < 0 || C <= 0 || h <= 0 || w <= 0 || n_tiles < 0) return fail(LC2IS_ERR_SHAPE, "bad shape%s");
    if (mode != LC2IS_BILINEAR && mode != LC2IS_BICUBIC) return fail(LC2IS_ERR_ARG, "bad mode%s");
    if (N == 0 || n_tiles == 0) return 0;
    if (!d_low || !d_desc) return fail(LC2IS_ERR_ARG, "null pointer%s");
    if (!d_labels && (d_confmat || d_per_image)) return fail(LC2IS_ERR_ARG, "statistics need labels%s");
    if (!d_pred && !d_confmat && !d_per_image) return fail(LC2IS_ERR_ARG, "no output requested%s");
    RaggedParams P;
    P.low = d_low; P.desc = (const long long*)d_desc; P.labels = (const long long*)d_labels;
    P.confmat = (unsigned long long*)d_confmat; P.per_image = (unsigned long long*)d_per_image;
    P.pred = (long long*)d_pred; P.N = N; P.C = C; P.h = h; P.w = w; P.ntiles = n_tiles;
    const size_t smem = d_per_image ? (size_t)3 * C * sizeof(int) : 0;
    if (smem > 48 * 1024) return fail(LC2IS_ERR_UNSUPPORTED, "too many classes for the per-image statistics%s");
    long long grid = (long long)sm_count() * 4;
    if (grid > n_tiles) grid = n_tiles;
    if (mode == LC2IS_BILINEAR) k3_ragged_kernel<0><<<(unsigned)grid, RG_THREADS, smem, (cudaStream_t)stream>>>(P);
    else k3_ragged_kernel<1><<<(unsigned)grid, RG_THREADS, smem, (cudaStream_t)stream>>>(P);
    LC2IS_CHECK_LAUNCH("k3_ragged_kernel");
    return 0;
}
