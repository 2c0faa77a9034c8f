// Producer of the projection GEMM: bicubic x4 upsample of the decoder's TOKEN-major feature map, written directly as the
// row-major [B*P, C] operand of TextToPatch.visual (bf16 for the tcgen05 GEMM, or fp32).
//
// Replaces reference model/model.py:42-44
//     dec_v = rearrange(dec_v, "b (h w) c -> b c h w", h=H)
//     dec_v = F.interpolate(input=dec_v, mode="bicubic", scale_factor=4)
//     dec_v = rearrange(dec_v, "b c h w -> b (h w) c", h=self.out_size)
// (two layout changes and an fp32 [B,C,4h,4w] intermediate) and their autograd backward.  In token-major layout the
// channels are contiguous, so every access is a coalesced vector and the upsample is "16 input tokens -> 16 output
// tokens" per GROUP: output rows 4k+2 .. 4k+5 all interpolate between input rows k-1 .. k+2 (ATen: src = 0.25*(dst+0.5)
// - 0.5, taps floor(src)-1 .. +2, index-clamped, A = -0.75), likewise the columns.
//
//   forward   one thread = one group row x 4 channels, walking the groups with the 4 x 4 token window in registers: 4 vector
//             loads per group, horizontal 4-tap chains, then vertical (ATen's order: x0*c0 + x1*c1 + x2*c2 + x3*c3 along x,
//             then along y), 16 vector stores.  HBM-bound on the output write.
//   backward  the transpose, separable, without atomics: T[cy][X] = sum_Y Wy(Y->cy) gy[Y][X] (fp32 scratch), then
//             gx[cy][cx] = sum_X Wx(X->cx) T[cy][X]; each pass reads its input ONCE (four sliding accumulators per thread;
//             taps clamped beyond the border fold onto the border cells).
#include "common.cuh"

namespace lc2is {

constexpr int BT_S = 4;
constexpr int BT_SEG = 8;                                  // groups per thread of the forward kernel

struct F4 { float x, y, z, w; };
__device__ __forceinline__ F4 f4_mul(F4 a, float s) { return {a.x * s, a.y * s, a.z * s, a.w * s}; }
__device__ __forceinline__ F4 f4_fma(F4 a, float s, F4 c) {
    return {fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z), fmaf(a.w, s, c.w)};
}
template <typename T> __device__ __forceinline__ F4 ld4(const T* p);
template <> __device__ __forceinline__ F4 ld4<float>(const float* p) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    return {v.x, v.y, v.z, v.w};
}
template <> __device__ __forceinline__ F4 ld4<__nv_bfloat16>(const __nv_bfloat16* p) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return {__uint_as_float(v.x << 16), __uint_as_float(v.x & 0xffff0000u), __uint_as_float(v.y << 16),
            __uint_as_float(v.y & 0xffff0000u)};
}
template <typename T> __device__ __forceinline__ void st4(T* p, F4 v);
template <> __device__ __forceinline__ void st4<float>(float* p, F4 v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v.x, v.y, v.z, v.w));
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, F4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 o;
    o.x = *reinterpret_cast<const unsigned*>(&lo);
    o.y = *reinterpret_cast<const unsigned*>(&hi);
    __stcs(reinterpret_cast<uint2*>(p), o);
}

// tap weights of output index 4k+2+r (r = 0..3) of group k: t = 0.125, 0.375, 0.625, 0.875 (exact in ATen's fp32 formula)
__device__ __forceinline__ void bt_coeffs(int r, float (&c)[4]) { cubic_coeffs(0.125f + 0.25f * (float)r, c); }

template <typename TI, typename TO>
__global__ void __launch_bounds__(256)
bicubic4_tokens_fwd_kernel(const TI* __restrict__ x, int B, int h, int w, int C, TO* __restrict__ y) {
    // one thread = BT_SEG consecutive groups of one group row (b, ky) x 4 channels: it walks them keeping the 4 x 4 window of
    // input tokens in registers and loads only the window's new column per step (an input token is loaded ~6 times instead of
    // 16); segments rather than whole rows: 128 registers = 16 warps per SM, and whole rows were 1.3 waves of long threads
    const int cv = C / 4;                                    // channel vectors per token
    const int nseg = (w + 1 + BT_SEG - 1) / BT_SEG;
    const long long total = (long long)B * (h + 1) * nseg * cv;
    const int H = BT_S * h, W = BT_S * w;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cv);
        long long g = idx / cv;
        const int kx0 = (int)(g % nseg) * BT_SEG - 1;
        g /= nseg;
        const int ky = (int)(g % (h + 1)) - 1;
        const int b = (int)(g / (h + 1));
        const int kx1 = min(w - 1, kx0 + BT_SEG - 1);
        const TI* xb = x + (size_t)b * h * w * C + (size_t)c4 * 4;
        size_t rowoff[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) rowoff[a] = (size_t)clampi(ky - 1 + a, 0, h - 1) * w * C;
        F4 t[4][4];                                          // t[a][bb]: input row a, window column bb (kx-1+bb, clamped)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int bb = 1; bb < 4; ++bb)                   // the window of group kx0 - 1; shifted by one column below
                t[a][bb] = ld4<TI>(xb + rowoff[a] + (size_t)clampi(kx0 - 2 + bb, 0, w - 1) * C);
#pragma unroll 1
        for (int kx = kx0; kx <= kx1; ++kx) {
            const size_t coloff = (size_t)clampi(kx + 2, 0, w - 1) * C;
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                t[a][0] = t[a][1]; t[a][1] = t[a][2]; t[a][2] = t[a][3];
                t[a][3] = ld4<TI>(xb + rowoff[a] + coloff);
            }
#pragma unroll
            for (int rx = 0; rx < 4; ++rx) {
                const int X = BT_S * kx + 2 + rx;
                if (X < 0 || X >= W) continue;
                float cx[4];
                bt_coeffs(rx, cx);
                F4 hx[4];                                    // the x-interpolated value of the four input rows at column X
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    F4 s = f4_mul(t[a][0], cx[0]);
#pragma unroll
                    for (int bb = 1; bb < 4; ++bb) s = f4_fma(t[a][bb], cx[bb], s);
                    hx[a] = s;
                }
#pragma unroll
                for (int ry = 0; ry < 4; ++ry) {
                    const int Y = BT_S * ky + 2 + ry;
                    if (Y < 0 || Y >= H) continue;
                    float cy[4];
                    bt_coeffs(ry, cy);
                    F4 s = f4_mul(hx[0], cy[0]);
#pragma unroll
                    for (int a = 1; a < 4; ++a) s = f4_fma(hx[a], cy[a], s);
                    st4<TO>(y + (((size_t)b * H + Y) * W + X) * C + (size_t)c4 * 4, s);
                }
            }
        }
    }
}

// One sweep of the transposed 1-D operator: out[cell] = sum_O W(O -> cell) in[O] over the n_out = 4 * n_in positions O, each
// read ONCE: the four accumulators are the (unclamped) cells k-1 .. k+2 of the current group k; when the group is done cell
// k-1 is complete and leaves (cells below 0 / above n_in-1 fold onto the border cells: ATen clamps the tap indices).
//   ld(O)     -> F4 value at position O          st(cell, F4) -> store the finished cell
template <typename LD, typename ST>
__device__ __forceinline__ void bt_sweep(int n_in, LD ld, ST st) {
    const int n_out = BT_S * n_in;
    F4 acc[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) acc[a] = {0.f, 0.f, 0.f, 0.f};
    F4 lo = {0.f, 0.f, 0.f, 0.f}, hi = {0.f, 0.f, 0.f, 0.f};
    auto add = [](F4 p, F4 q) { return F4{p.x + q.x, p.y + q.y, p.z + q.z, p.w + q.w}; };
    auto emit = [&](int u, F4 v) {                             // unclamped cell u is complete
        if (u <= 0) {
            lo = add(lo, v);
            if (u == 0 && n_in > 1) st(0, lo);
        } else if (u < n_in - 1) {
            st(u, v);
        } else {
            hi = add(hi, v);
        }
    };
#pragma unroll 1
    for (int k = -1; k < n_in; ++k) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int O = BT_S * k + 2 + r;
            if (O < 0 || O >= n_out) continue;
            float c[4];
            bt_coeffs(r, c);
            const F4 v = ld(O);
#pragma unroll
            for (int a = 0; a < 4; ++a) acc[a] = f4_fma(v, c[a], acc[a]);
        }
        emit(k - 1, acc[0]);
        acc[0] = acc[1]; acc[1] = acc[2]; acc[2] = acc[3]; acc[3] = {0.f, 0.f, 0.f, 0.f};
    }
    emit(n_in - 1, acc[0]); emit(n_in, acc[1]); emit(n_in + 1, acc[2]);
    if (n_in == 1) hi = add(hi, lo);
    st(n_in - 1, hi);
}

// pass 1 of the backward: T[b][cy][X][c] = sum_Y Wy(Y -> cy) gy[b][Y][X][c]; one thread = (b, X, 4 channels), all rows
template <typename TG>
__global__ void __launch_bounds__(256)
bicubic4_tokens_bwd_rows_kernel(const TG* __restrict__ gy, int B, int h, int w, int C, float* __restrict__ T) {
    const int cv = C / 4, W = BT_S * w, H = BT_S * h;
    const long long total = (long long)B * W * cv;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cv);
        const long long g = idx / cv;
        const int X = (int)(g % W), b = (int)(g / W);
        const TG* src = gy + ((size_t)b * H * W + X) * C + (size_t)c4 * 4;
        float* dst = T + ((size_t)b * h * W + X) * C + (size_t)c4 * 4;
        bt_sweep(h, [&](int Y) { return ld4<TG>(src + (size_t)Y * W * C); },
                 [&](int cy, F4 v) { *reinterpret_cast<float4*>(dst + (size_t)cy * W * C) = make_float4(v.x, v.y, v.z, v.w); });
    }
}

// pass 2: gx[b][cy][cx][c] = sum_X Wx(X -> cx) T[b][cy][X][c]; one thread = (b, cy, 4 channels), all columns
template <typename TO>
__global__ void __launch_bounds__(256)
bicubic4_tokens_bwd_cols_kernel(const float* __restrict__ T, int B, int h, int w, int C, TO* __restrict__ gx) {
    const int cv = C / 4, W = BT_S * w;
    const long long total = (long long)B * h * cv;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(idx % cv);
        const long long g = idx / cv;                          // g = b * h + cy
        const float* src = T + (size_t)g * W * C + (size_t)c4 * 4;
        TO* dst = gx + (size_t)g * w * C + (size_t)c4 * 4;
        bt_sweep(w, [&](int X) { const float4 v = __ldcs(reinterpret_cast<const float4*>(src + (size_t)X * C)); return F4{v.x, v.y, v.z, v.w}; },
                 [&](int cx, F4 v) { st4<TO>(dst + (size_t)cx * C, v); });
    }
}

static unsigned bt_grid(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(blocks > cap ? cap : (blocks < 1 ? 1 : blocks));
}

}  // namespace lc2is

using namespace lc2is;

static int bt_check(const void* a, const void* b, int B, int h, int w, int C, int dt_a, int dt_b) {
    if (B < 0 || h <= 0 || w <= 0 || C <= 0) return fail(LC2IS_ERR_SHAPE, "bad shape%s");
    if (C % 4) return fail(LC2IS_ERR_SHAPE, "C must be a multiple of 4 (got %s%lld)", "", C);
    if (!a || !b) return fail(LC2IS_ERR_ARG, "null pointer%s");
    if (((uintptr_t)a | (uintptr_t)b) % 16) return fail(LC2IS_ERR_ARG, "pointers must be 16-byte aligned%s");
    if ((dt_a != LC2IS_F32 && dt_a != LC2IS_BF16) || (dt_b != LC2IS_F32 && dt_b != LC2IS_BF16))
        return fail(LC2IS_ERR_ARG, "dtype must be LC2IS_F32 or LC2IS_BF16%s");
    return 0;
}

extern "C" int lc2is_bicubic4_tokens_fwd(const void* d_x, int x_dtype, int B, int h, int w, int C, void* d_y, int y_dtype,
                                         lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = bt_check(d_x, d_y, B, h, w, C, x_dtype, y_dtype)) return e;
    if (B == 0) return 0;
    const long long total = (long long)B * (h + 1) * ((w + 1 + BT_SEG - 1) / BT_SEG) * (C / 4);
    const unsigned grid = bt_grid(total);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == LC2IS_F32 && y_dtype == LC2IS_BF16)
        bicubic4_tokens_fwd_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>((const float*)d_x, B, h, w, C, (__nv_bfloat16*)d_y);
    else if (x_dtype == LC2IS_F32)
        bicubic4_tokens_fwd_kernel<float, float><<<grid, 256, 0, st>>>((const float*)d_x, B, h, w, C, (float*)d_y);
    else if (y_dtype == LC2IS_BF16)
        bicubic4_tokens_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)d_x, B, h, w, C,
                                                                                       (__nv_bfloat16*)d_y);
    else
        bicubic4_tokens_fwd_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)d_x, B, h, w, C, (float*)d_y);
    LC2IS_CHECK_LAUNCH("bicubic4_tokens_fwd_kernel");
    return 0;
}

extern "C" int64_t lc2is_bicubic4_tokens_bwd_workspace(int B, int h, int w, int C) {
    return (int64_t)B * h * (BT_S * w) * C * 4;
}

extern "C" int lc2is_bicubic4_tokens_bwd(const void* d_gy, int gy_dtype, int B, int h, int w, int C, void* d_gx, int gx_dtype,
                                         void* d_ws, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (int e = bt_check(d_gy, d_gx, B, h, w, C, gy_dtype, gx_dtype)) return e;
    if (!d_ws || (uintptr_t)d_ws % 16) return fail(LC2IS_ERR_ARG, "workspace must be a 16-byte aligned device pointer%s");
    if (B == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    float* T = (float*)d_ws;
    const long long t1 = (long long)B * (BT_S * w) * (C / 4), t2 = (long long)B * h * (C / 4);
    if (gy_dtype == LC2IS_BF16)
        bicubic4_tokens_bwd_rows_kernel<__nv_bfloat16><<<bt_grid(t1), 256, 0, st>>>((const __nv_bfloat16*)d_gy, B, h, w, C, T);
    else
        bicubic4_tokens_bwd_rows_kernel<float><<<bt_grid(t1), 256, 0, st>>>((const float*)d_gy, B, h, w, C, T);
    LC2IS_CHECK_LAUNCH("bicubic4_tokens_bwd_rows_kernel");
    if (gx_dtype == LC2IS_BF16)
        bicubic4_tokens_bwd_cols_kernel<__nv_bfloat16><<<bt_grid(t2), 256, 0, st>>>(T, B, h, w, C, (__nv_bfloat16*)d_gx);
    else
        bicubic4_tokens_bwd_cols_kernel<float><<<bt_grid(t2), 256, 0, st>>>(T, B, h, w, C, (float*)d_gx);
    LC2IS_CHECK_LAUNCH("bicubic4_tokens_bwd_cols_kernel");
    return 0;
}
