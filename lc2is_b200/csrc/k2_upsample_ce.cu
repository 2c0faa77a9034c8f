// K2: bilinear upsample + softmax cross-entropy, forward AND backward in one pass.
//
// Replaces reference model/loss.py:19-20 (AuxiliaryLoss: F.interpolate(bilinear, size=H) +
// CrossEntropyLoss) / model/final.py:44 + engine.py:94, and their autograd backward
// (engine.py:100).  The upsampled [B,C,H,W] tensor (2.5 GB at B=16, C=150, 512^2) never
// exists: each CTA stages the low-resolution logits of a pixel tile in shared memory,
// evaluates the softmax of every upsampled pixel in registers and scatters
// (softmax - onehot) * g back through the four bilinear taps into a shared-memory gradient
// tile that is flushed once.
//
// Fast path (power-of-two scale s >= 4): one thread owns a 4x4 pixel block whose 16 pixels
// share the same 2x2 taps (common.cuh: BlockGeom).  Inside the block the upsampled logit is a
// bilinear polynomial  l(i,j) = l00 + i*P + j*Q + i*j*T, so
//     exp(l(i,j) - M) = E00 * p^i * q^j * t^(i*j)        (4 ex2 per class per 16 pixels,
// the rest are multiplies), with M = max over classes of the block's four taps (an upper
// bound of every pixel's max - a valid softmax shift).  If a pixel's sum under/overflows
// with that shared shift (logits spanning > ~80 within one cell) the thread falls back to
// the exact per-pixel path.  The scatter uses the moments sum(g), sum(i g), sum(j g),
// sum(ij g) instead of 64 weighted adds.
//
// Generic path (any size): one pixel per thread, taps from global, global atomics.
#include "common.cuh"

namespace lc2is {

constexpr float LOG2E = 1.4426950408889634f;
constexpr float LN2 = 0.6931471805599453f;
constexpr float S_MIN = 1e-30f;       // below this the shared shift lost precision -> slow path

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// count valid labels (the 'mean' denominator)
__global__ void k2_count_valid_kernel(const long long* __restrict__ labels, long long n, long long ignore,
                                      unsigned long long* __restrict__ out) {
    long long cnt = 0;
    const long long stride = (long long)gridDim.x * blockDim.x;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // 2 labels (16 B) per load when aligned
    const long long n2 = n / 2;
    const longlong2* l2 = reinterpret_cast<const longlong2*>(labels);
    for (long long k = i; k < n2; k += stride) {
        longlong2 v = __ldg(l2 + k);
        cnt += (v.x != ignore) + (v.y != ignore);
    }
    if (i == 0 && (n & 1)) cnt += labels[n - 1] != ignore;
    int c = (int)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    __shared__ int ws[32];
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = threadIdx.x < (blockDim.x >> 5) ? ws[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && v) atomicAdd(out, (unsigned long long)v);
    }
}

// ---------------------------------------------------------------------------------------------
// Geometry of the fast path.  A GROUP is the set of bps x bps 4x4-pixel blocks (bps = s/4) that
// interpolate between the same 2x2 source cells: group (gy,gx) has taps (ky,kx) = (gy-1, gx-1)
// .. (ky+1, kx+1), index-clamped to the image.  There are (h+1) x (w+1) groups per image.  The
// bps*bps threads of a group are consecutive lanes of one warp, so their four tap contributions
// are summed with a butterfly of warp shuffles and leave the SM as four L2 float reductions.
struct K2Params {
    const float* low;            // [B,C,h,w]
    const long long* labels;     // [B,H,W]
    float* grad_low;             // [B,C,h,w] (pre-zeroed; accumulated with global reductions) or null
    double* loss_sum;
    const float* grad_scale;     // device scalar or null
    long long ignore_index;
    int B, C, h, w, H, W;
    int s, off, q;                // scale, pixel offset of the block grid, block offset of group 0
    int tgy, tgx;                 // groups per CTA
    float rs;
};

__device__ __forceinline__ int clampi2(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// Sum (a,b,c,d) over the LPG consecutive lanes of a group.  After the call lane u of the group
// holds in `a` the total of value number (u * 4 / LPG) [LPG >= 4], i.e. the first quarter of the
// lanes hold sum(a), the second sum(b), ... ; for LPG == 1 nothing happens.
template <int LPG>
__device__ __forceinline__ void group_reduce4(float& a, float& b, float& c, float& d, int u) {
    if constexpr (LPG >= 4) {
        // step 1: halves exchange pairs -> lower half keeps (a,b), upper half keeps (c,d)
        const bool up = u & (LPG / 2);
        float s0 = up ? a : c, s1 = up ? b : d;          // what this lane gives away
        float k0 = up ? c : a, k1 = up ? d : b;          // what it keeps
        k0 += __shfl_xor_sync(0xffffffffu, s0, LPG / 2);
        k1 += __shfl_xor_sync(0xffffffffu, s1, LPG / 2);
        // step 2: quarters -> one value per lane
        const bool up2 = u & (LPG / 4);
        float g = up2 ? k0 : k1;
        float k = up2 ? k1 : k0;
        k += __shfl_xor_sync(0xffffffffu, g, LPG / 4);
#pragma unroll
        for (int o = LPG / 8; o > 0; o >>= 1) k += __shfl_xor_sync(0xffffffffu, k, o);
        a = k;
    }
}

template <int BPS>
__global__ void __launch_bounds__(128)
k2_fast_kernel(const K2Params P) {
    constexpr int LPG = BPS * BPS;                          // lanes (threads) per group
    extern __shared__ float smem[];
    const int C = P.C;
    const int ncx = P.tgx + 1, ncy = P.tgy + 1;
    const int ncell = ncy * ncx;
    const int cs = ncell;                                   // class stride in the tile
    const int nthr = blockDim.x;
    float* st = smem;                                       // source tile  [C][ncy][ncx]
    float* cellmax = smem + (size_t)C * cs;                 // [ncell]
    __shared__ float red[4];

    const int n = blockIdx.z;
    const int GY0 = blockIdx.y * P.tgy, GX0 = blockIdx.x * P.tgx;
    const int cy0 = GY0 - 1, cx0 = GX0 - 1;                 // first cell of the tile (may be -1)
    const float* lowb = P.low + (size_t)n * C * P.h * P.w;

    // ---- stage the source tile ------------------------------------------------------------------
    for (int idx = threadIdx.x; idx < C * ncell; idx += nthr) {
        int c = idx / ncell, r = idx - c * ncell;
        int i = r / ncx, j = r - i * ncx;
        int gy = cy0 + i, gx = cx0 + j;
        if (gy >= 0 && gy < P.h && gx >= 0 && gx < P.w)
            cp_async4(st + idx, lowb + ((size_t)c * P.h + gy) * P.w + gx);
        else
            st[idx] = -INFINITY;
    }
    cp_async_wait_all();
    __syncthreads();
    for (int r = threadIdx.x; r < ncell; r += nthr) {
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) m = fmaxf(m, st[c * cs + r]);
        cellmax[r] = m;
    }
    __syncthreads();

    // ---- per-thread block setup -------------------------------------------------------------------
    const int g = threadIdx.x / LPG, u = threadIdx.x % LPG;
    const int tgy = g / P.tgx, tgx = g - tgy * P.tgx;
    const int uy = u / BPS, ux = u % BPS;
    const int ky = GY0 + tgy - 1, kx = GX0 + tgx - 1;       // top-left tap of the group (-1 .. h-1)
    const int by = ky * BPS + P.q + uy, bx = kx * BPS + P.q + ux;
    const int y0 = 4 * by - P.off, x0 = 4 * bx - P.off;
    const float gs = P.grad_scale ? __ldg(P.grad_scale) : 1.f;
    const float rs = P.rs;
    const bool group_in = ky < P.h && kx < P.w;             // group exists in this image
    int lab[16];
    bool any = false;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int y = y0 + i, x = x0 + j;
            int l = -1;
            if (group_in && y >= 0 && y < P.H && x >= 0 && x < P.W) {
                long long t = __ldg(P.labels + ((size_t)n * P.H + y) * P.W + x);
                if (t != P.ignore_index && t >= 0 && t < C) l = (int)t;
            }
            lab[i * 4 + j] = l;
            any |= l >= 0;
        }
    // clamped tap cells (global) and their offsets in the tile
    const int Ya = clampi2(ky, 0, P.h - 1), Yb = clampi2(ky + 1, 0, P.h - 1);
    const int Xa = clampi2(kx, 0, P.w - 1), Xb = clampi2(kx + 1, 0, P.w - 1);
    int oa = 0, ob = 0, oc = 0, od = 0;
    float ly0 = 0.f, lx0 = 0.f, M = 0.f;
    if (group_in) {
        oa = (Ya - cy0) * ncx + (Xa - cx0); ob = (Ya - cy0) * ncx + (Xb - cx0);
        oc = (Yb - cy0) * ncx + (Xa - cx0); od = (Yb - cy0) * ncx + (Xb - cx0);
        // lambda of the block's first row / column ((y0+0.5)*rs-0.5-ky is exact for power-of-2 s)
        ly0 = ((float)y0 + 0.5f) * rs - 0.5f - (float)ky;
        lx0 = ((float)x0 + 0.5f) * rs - 0.5f - (float)kx;
        M = fmaxf(fmaxf(cellmax[oa], cellmax[ob]), fmaxf(cellmax[oc], cellmax[od]));
    }
    const float Mk = M * LOG2E;
    const float k1 = rs * LOG2E, k2 = rs * rs * LOG2E;

    // ---- pass A: S(i,j) = sum_c exp(l_c(i,j) - M) ----------------------------------------------------
    float S[16];                      // becomes U = g / S after pass A
    bool slow = false;
    float loss = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) S[i] = 0.f;
    if (any) {
#pragma unroll 2
        for (int c = 0; c < C; ++c) {
            const float* p = st + c * cs;
            const float a = p[oa], b = p[ob], cc = p[oc], d = p[od];
            const float da = cc - a, db = d - b, dd = db - da;
            const float L0 = fmaf(ly0, da, a), R0 = fmaf(ly0, db, b);
            const float rl = R0 - L0;
            const float l00 = fmaf(lx0, rl, L0);
            const float q0 = ex2f(rl * k1);
            const float pp = ex2f(fmaf(lx0, dd, da) * k1);
            const float tt = ex2f(dd * k2);
            float e0 = ex2f(fmaf(l00, LOG2E, -Mk));
            float qi = q0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float e = e0;
                S[i * 4 + 0] += e;
                e *= qi; S[i * 4 + 1] += e;
                e *= qi; S[i * 4 + 2] += e;
                e *= qi; S[i * 4 + 3] += e;
                e0 *= pp;
                qi *= tt;
            }
        }
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (lab[i] >= 0) slow |= !((S[i] >= S_MIN) && (S[i] <= 3.0e38f));
    }
    // The exact fallback is taken by the whole warp (the scatter pass below is warp-collective).
    slow = __any_sync(0xffffffffu, slow);
    float Mp[16];                     // per-pixel softmax shift (only differs from M on the slow path)
#pragma unroll
    for (int i = 0; i < 16; ++i) Mp[i] = M;
    if (slow && any) {
        // exact per-pixel max and sum (logits spanning > ~80 inside one cell, inf/NaN, ...)
#pragma unroll 1
        for (int pix = 0; pix < 16; ++pix) {
            const float ly = ly0 + (float)(pix >> 2) * rs, lx = lx0 + (float)(pix & 3) * rs;
            float m = -INFINITY;
            for (int c = 0; c < C; ++c) {
                const float* p = st + c * cs;
                float L = fmaf(ly, p[oc] - p[oa], p[oa]), R = fmaf(ly, p[od] - p[ob], p[ob]);
                m = fmaxf(m, fmaf(lx, R - L, L));
            }
            float sum = 0.f;
            for (int c = 0; c < C; ++c) {
                const float* p = st + c * cs;
                float L = fmaf(ly, p[oc] - p[oa], p[oa]), R = fmaf(ly, p[od] - p[ob], p[ob]);
                sum += ex2f((fmaf(lx, R - L, L) - m) * LOG2E);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (i == pix) { Mp[i] = m; S[i] = sum; }
        }
    }
    // ---- loss and per-pixel gradient scale U = g / S --------------------------------------------------
    if (any) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = lab[i * 4 + j];
                float uu = 0.f;
                if (t >= 0) {
                    const float ly = ly0 + (float)i * rs, lx = lx0 + (float)j * rs;
                    const float* p = st + t * cs;
                    const float L = fmaf(ly, p[oc] - p[oa], p[oa]), R = fmaf(ly, p[od] - p[ob], p[ob]);
                    const float lt = fmaf(lx, R - L, L);
                    loss += logf(S[i * 4 + j]) + Mp[i * 4 + j] - lt;
                    uu = gs / S[i * 4 + j];
                }
                S[i * 4 + j] = uu;
            }
    }

    // ---- pass B: scatter (softmax - onehot) * g through the taps ---------------------------------------
    // Warp-collective: every lane runs the class loop (lanes without valid pixels contribute zeros).
    float* gb = P.grad_low ? P.grad_low + (size_t)n * C * P.h * P.w : nullptr;
    const bool warp_any = __any_sync(0xffffffffu, any);
    if (gb && warp_any) {
        // which of the four totals this lane writes after the butterfly, and where
        const int role = LPG >= 4 ? (u * 4) / LPG : 0;
        const bool writer = LPG >= 4 ? (u % (LPG / 4 > 0 ? LPG / 4 : 1)) == 0 : true;
        const int cellY = (role & 2) ? Yb : Ya, cellX = (role & 1) ? Xb : Xa;
        float* gcell = gb + (size_t)cellY * P.w + cellX;
        const size_t plane = (size_t)P.h * P.w;
        int cur = 0x7fffffff;                              // smallest label among this thread's valid pixels
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if (lab[i] >= 0) cur = min(cur, lab[i]);
        for (int c = 0; c < C; ++c) {
            float A = 0.f, Bv = 0.f, Cv = 0.f, Dv = 0.f;
            if (any) {
                const float* p = st + c * cs;
                const float a = p[oa], b = p[ob], cc = p[oc], d = p[od];
                float G, Gx, Gy, Gxy;
                if (!slow) {
                    const float da = cc - a, db = d - b, dd = db - da;
                    const float L0 = fmaf(ly0, da, a), R0 = fmaf(ly0, db, b);
                    const float rl = R0 - L0;
                    const float l00 = fmaf(lx0, rl, L0);
                    const float q0 = ex2f(rl * k1);
                    const float pp = ex2f(fmaf(lx0, dd, da) * k1);
                    const float tt = ex2f(dd * k2);
                    float e0 = ex2f(fmaf(l00, LOG2E, -Mk));
                    float qi = q0;
                    float r[4], rx[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float e = e0;
                        const float g0 = e * S[i * 4 + 0];
                        e *= qi; const float g1 = e * S[i * 4 + 1];
                        e *= qi; const float g2 = e * S[i * 4 + 2];
                        e *= qi; const float g3 = e * S[i * 4 + 3];
                        r[i] = (g0 + g1) + (g2 + g3);
                        rx[i] = fmaf(3.f, g3, fmaf(2.f, g2, g1));
                        e0 *= pp;
                        qi *= tt;
                    }
                    G = (r[0] + r[1]) + (r[2] + r[3]);
                    Gy = fmaf(3.f, r[3], fmaf(2.f, r[2], r[1]));
                    Gx = (rx[0] + rx[1]) + (rx[2] + rx[3]);
                    Gxy = fmaf(3.f, rx[3], fmaf(2.f, rx[2], rx[1]));
                } else {
                    G = Gx = Gy = Gxy = 0.f;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float ly = ly0 + (float)i * rs, lx = lx0 + (float)j * rs;
                            const float L = fmaf(ly, cc - a, a), R = fmaf(ly, d - b, b);
                            const float gg = ex2f((fmaf(lx, R - L, L) - Mp[i * 4 + j]) * LOG2E) * S[i * 4 + j];
                            G += gg; Gx += (float)j * gg; Gy += (float)i * gg; Gxy += (float)(i * j) * gg;
                        }
                }
                const float X = fmaf(rs, Gx, lx0 * G);                        // sum lambda_x g
                const float Y = fmaf(rs, Gy, ly0 * G);                        // sum lambda_y g
                const float XY = fmaf(ly0, X, rs * fmaf(rs, Gxy, lx0 * Gy));  // sum lambda_x lambda_y g
                A = (G - X) - (Y - XY); Bv = X - XY; Cv = Y - XY; Dv = XY;
                // - g * onehot for the pixels whose target is this class.  `cur` is the smallest
                // not-yet-handled label of this thread: one compare per class, the 16-way scan only on a hit.
                if (c == cur) {
                    int nxt = 0x7fffffff;
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int t = lab[i * 4 + j];
                            if (t == c) {
                                const float ly = ly0 + (float)i * rs, lx = lx0 + (float)j * rs;
                                A -= gs * (1.f - ly) * (1.f - lx); Bv -= gs * (1.f - ly) * lx;
                                Cv -= gs * ly * (1.f - lx);        Dv -= gs * ly * lx;
                            } else if (t > c) {
                                nxt = min(nxt, t);
                            }
                        }
                    cur = nxt;
                }
            }
            group_reduce4<LPG>(A, Bv, Cv, Dv, u);
            if constexpr (LPG >= 4) {
                if (writer && A != 0.f) atomicAdd(gcell + (size_t)c * plane, A);
            } else {
                float* gp = gb + (size_t)c * plane;
                if (any) {
                    if (A != 0.f) atomicAdd(gp + (size_t)Ya * P.w + Xa, A);
                    if (Bv != 0.f) atomicAdd(gp + (size_t)Ya * P.w + Xb, Bv);
                    if (Cv != 0.f) atomicAdd(gp + (size_t)Yb * P.w + Xa, Cv);
                    if (Dv != 0.f) atomicAdd(gp + (size_t)Yb * P.w + Xb, Dv);
                }
            }
        }
    }
    // ---- loss reduction: warp -> CTA -> one double atomic ----------------------------------------------
    loss = warp_sum(loss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (nthr >> 5); ++i) t += red[i];
        if (t != 0.f) atomicAdd(P.loss_sum, (double)t);
    }
}

// ---------------------------------------------------------------------------------------------
// generic path: any (h,w)->(H,W); one pixel per thread; exact 3-pass softmax; global atomics.
__global__ void __launch_bounds__(256)
k2_generic_kernel(const float* __restrict__ low, const long long* __restrict__ labels,
                  float* __restrict__ grad_low, double* __restrict__ loss_sum,
                  const float* __restrict__ grad_scale, long long ignore_index,
                  int B, int C, int h, int w, int H, int W, float scale_y, float scale_x) {
    const long long HW = (long long)H * W;
    const long long total = (long long)B * HW;
    const float gs = grad_scale ? __ldg(grad_scale) : 1.f;
    float loss = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total;
         p += (long long)gridDim.x * blockDim.x) {
        const int n = (int)(p / HW);
        const long long r = p - (long long)n * HW;
        const int y = (int)(r / W), x = (int)(r - (long long)y * W);
        const long long tl = labels[p];
        if (tl == ignore_index || tl < 0 || tl >= C) continue;
        const int t = (int)tl;
        float sy = scale_y * ((float)y + 0.5f) - 0.5f; if (sy < 0.f) sy = 0.f;
        float sx = scale_x * ((float)x + 0.5f) - 0.5f; if (sx < 0.f) sx = 0.f;
        int ky = (int)sy; if (ky > h - 1) ky = h - 1;
        int kx = (int)sx; if (kx > w - 1) kx = w - 1;
        float ly = fminf(fmaxf(sy - (float)ky, 0.f), 1.f), lx = fminf(fmaxf(sx - (float)kx, 0.f), 1.f);
        const int ky1 = ky + 1 < h ? ky + 1 : h - 1, kx1 = kx + 1 < w ? kx + 1 : w - 1;
        const int oa = ky * w + kx, ob = ky * w + kx1, oc = ky1 * w + kx, od = ky1 * w + kx1;
        const float wa = (1.f - ly) * (1.f - lx), wb = (1.f - ly) * lx, wc = ly * (1.f - lx), wd = ly * lx;
        const float* base = low + (size_t)n * C * h * w;
        float m = -INFINITY;
        for (int c = 0; c < C; ++c) {
            const float* q = base + (size_t)c * h * w;
            float l = (1.f - ly) * ((1.f - lx) * __ldg(q + oa) + lx * __ldg(q + ob)) +
                      ly * ((1.f - lx) * __ldg(q + oc) + lx * __ldg(q + od));
            m = fmaxf(m, l);
        }
        float S = 0.f, lt = 0.f;
        for (int c = 0; c < C; ++c) {
            const float* q = base + (size_t)c * h * w;
            float l = (1.f - ly) * ((1.f - lx) * __ldg(q + oa) + lx * __ldg(q + ob)) +
                      ly * ((1.f - lx) * __ldg(q + oc) + lx * __ldg(q + od));
            S += ex2f((l - m) * LOG2E);
            if (c == t) lt = l;
        }
        loss += lg2f(S) * LN2 + m - lt;
        if (grad_low) {
            const float u = gs / S;
            float* gb = grad_low + (size_t)n * C * h * w;
            for (int c = 0; c < C; ++c) {
                const float* q = base + (size_t)c * h * w;
                float l = (1.f - ly) * ((1.f - lx) * __ldg(q + oa) + lx * __ldg(q + ob)) +
                          ly * ((1.f - lx) * __ldg(q + oc) + lx * __ldg(q + od));
                float g = ex2f((l - m) * LOG2E) * u - (c == t ? gs : 0.f);
                float* gp = gb + (size_t)c * h * w;
                atomicAdd(gp + oa, g * wa); atomicAdd(gp + ob, g * wb);
                atomicAdd(gp + oc, g * wc); atomicAdd(gp + od, g * wd);
            }
        }
    }
    loss = warp_sum(loss);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = loss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)(blockDim.x >> 5); ++i) t += red[i];
        if (t != 0.f) atomicAdd(loss_sum, (double)t);
    }
}

// fp32 [B,C,hw] -> bf16 [B,C_pad,hw], zero pad rows
__global__ void k2_grad_to_bf16_kernel(const float* __restrict__ g, int C, int C_pad, long long hw,
                                       __nv_bfloat16* __restrict__ out, long long total) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        long long p = i % hw;
        long long bc = i / hw;
        int c = (int)(bc % C_pad);
        long long b = bc / C_pad;
        float v = c < C ? g[((size_t)b * C + c) * hw + p] : 0.f;
        out[i] = __float2bfloat16(v);
    }
}

}  // namespace lc2is

using namespace lc2is;

extern "C" int lc2is_count_valid(const int64_t* d_labels, int64_t n, int64_t ignore_index,
                                 int64_t* d_n_valid, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (n < 0) return fail(LC2IS_ERR_SHAPE, "negative n%s");
    if (n == 0) return 0;
    if (!d_labels || !d_n_valid) return fail(LC2IS_ERR_ARG, "null pointer%s");
    if ((uintptr_t)d_labels % 16) return fail(LC2IS_ERR_ARG, "labels must be 16-byte aligned%s");
    long long blocks = (n / 2 + 255) / 256;
    long long cap = (long long)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k2_count_valid_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        (const long long*)d_labels, n, ignore_index, (unsigned long long*)d_n_valid);
    LC2IS_CHECK_LAUNCH("k2_count_valid_kernel");
    return 0;
}

extern "C" int lc2is_grad_to_bf16(const float* d_grad, int B, int C, int hw, void* d_grad_bf16,
                                  lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (!d_grad || !d_grad_bf16) return fail(LC2IS_ERR_ARG, "null pointer%s");
    const int C_pad = class_pad(C);
    long long total = (long long)B * C_pad * hw;
    if (total == 0) return 0;
    long long blocks = (total + 255) / 256;
    long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    k2_grad_to_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
        d_grad, C, C_pad, hw, (__nv_bfloat16*)d_grad_bf16, total);
    LC2IS_CHECK_LAUNCH("k2_grad_to_bf16_kernel");
    return 0;
}

extern "C" int lc2is_upsample_ce_fwd_bwd(const float* d_low, const int64_t* d_labels,
                                         int B, int C, int h, int w, int H, int W,
                                         int64_t ignore_index, const float* d_grad_scale,
                                         double* d_loss_sum, float* d_grad_low, void* d_grad_low_bf16,
                                         lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (B < 0 || C <= 0 || h <= 0 || w <= 0 || H <= 0 || W <= 0) return fail(LC2IS_ERR_SHAPE, "bad shape%s");
    if (B == 0) return 0;
    if (!d_low || !d_labels || !d_loss_sum) return fail(LC2IS_ERR_ARG, "null pointer%s");
    if (d_grad_low_bf16 && !d_grad_low)
        return fail(LC2IS_ERR_ARG, "grad_low_bf16 needs the fp32 grad_low buffer as well%s");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (d_grad_low) LC2IS_CUDA(cudaMemsetAsync(d_grad_low, 0, (size_t)B * C * h * w * sizeof(float), st));

    int s = 0;
    bool fast = fast_scale(h, w, H, W, &s) && s <= 16;
    K2Params P;
    size_t smem = 0;
    if (fast) {
        BlockGeom g = make_geom(H, W, s);
        P.low = d_low; P.labels = (const long long*)d_labels; P.grad_low = d_grad_low;
        P.loss_sum = d_loss_sum; P.grad_scale = d_grad_scale; P.ignore_index = ignore_index;
        P.B = B; P.C = C; P.h = h; P.w = w; P.H = H; P.W = W;
        P.s = s; P.off = g.off; P.rs = g.rs;
        const int bps = s / 4;
        P.q = (g.off + s / 2) / 4;
        // 128 threads = tgy*tgx groups of bps*bps threads; tile = 32 x 64 pixels for every scale
        P.tgy = 8 / bps; P.tgx = 16 / bps;
        smem = ((size_t)C + 1) * (P.tgy + 1) * (P.tgx + 1) * sizeof(float);
        if (smem > 110 * 1024) fast = false;
    }
    if (fast) {
        dim3 grid((w + 1 + P.tgx - 1) / P.tgx, (h + 1 + P.tgy - 1) / P.tgy, B);
        auto launch = [&](auto kernel) -> int {
            LC2IS_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<grid, 128, smem, st>>>(P);
            return 0;
        };
        int e = s == 4 ? launch(k2_fast_kernel<1>) : s == 8 ? launch(k2_fast_kernel<2>) : launch(k2_fast_kernel<4>);
        if (e) return e;
        LC2IS_CHECK_LAUNCH("k2_fast_kernel");
    } else {
        const float sy = (float)h / (float)H, sx = (float)w / (float)W;
        long long total = (long long)B * H * W;
        long long blocks = (total + 255) / 256;
        long long cap = (long long)sm_count() * 8;
        if (blocks > cap) blocks = cap;
        k2_generic_kernel<<<(unsigned)blocks, 256, 0, st>>>(d_low, (const long long*)d_labels, d_grad_low,
                                                             d_loss_sum, d_grad_scale, ignore_index,
                                                             B, C, h, w, H, W, sy, sx);
        LC2IS_CHECK_LAUNCH("k2_generic_kernel");
    }
    if (d_grad_low_bf16) return lc2is_grad_to_bf16(d_grad_low, B, C, h * w, d_grad_low_bf16, stream);
    return 0;
}
