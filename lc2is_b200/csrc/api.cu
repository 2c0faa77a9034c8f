// C-ABI glue: error text, device probing, scalar helpers and the whole-step-from-host entry.
#include "common.cuh"

namespace lc2is {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

static int g_sm_count = 0;

int ensure_device() {
    static int state = 0;   // 0 unknown, 1 ok, -1 none
    if (state == 1) return 0;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        state = -1;
        return fail(LC2IS_ERR_NODEVICE, "no CUDA device available: lc2is_b200 has no CPU fallback%s");
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (p.major != 10) {
        snprintf(g_err, sizeof(g_err), "device '%s' is sm_%d%d; this library is built for sm_100a (B200) only",
                 p.name, p.major, p.minor);
        return LC2IS_ERR_NODEVICE;
    }
    g_sm_count = p.multiProcessorCount;
    state = 1;
    return 0;
}
int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

// scale = mult / max(n_valid, 1)   (n_valid == 0 -> scale = 0: no valid pixel, no gradient)
__global__ void mean_scale_kernel(const long long* n_valid, float mult, float* out) {
    long long n = *n_valid;
    *out = n > 0 ? mult / (float)n : 0.f;
}
// loss = loss_sum / n_valid   (0/0 -> NaN, as torch's CrossEntropyLoss('mean') with no valid target)
__global__ void finalize_loss_kernel(const double* loss_sum, const long long* n_valid, float* out) {
    *out = (float)(*loss_sum / (double)(*n_valid));
}

}  // namespace lc2is

using namespace lc2is;

extern "C" const char* lc2is_last_error(void) { return g_err; }
extern "C" int lc2is_abi_version(void) { return 1; }
extern "C" int64_t lc2is_launch_count(void) { return g_launches.load(); }

extern "C" int lc2is_mean_scale(const int64_t* d_n_valid, float mult, float* d_scale, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (!d_n_valid || !d_scale) return fail(LC2IS_ERR_ARG, "null pointer%s");
    mean_scale_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const long long*)d_n_valid, mult, d_scale);
    LC2IS_CHECK_LAUNCH("mean_scale_kernel");
    return 0;
}

extern "C" int lc2is_finalize_loss(const double* d_loss_sum, const int64_t* d_n_valid, float* d_loss,
                                   lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (!d_loss_sum || !d_n_valid || !d_loss) return fail(LC2IS_ERR_ARG, "null pointer%s");
    finalize_loss_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(d_loss_sum, (const long long*)d_n_valid, d_loss);
    LC2IS_CHECK_LAUNCH("finalize_loss_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// whole step from host buffers
namespace {
struct StepWs {
    size_t v_in, labels, packed, t_in, t_hat, inv_t, v_hat, inv_v, logits, grad_low, grad_v, grad_t,
        bwd_ws, confmat, scalars, total;
};
inline size_t al(size_t x) { return (x + 255) / 256 * 256; }
StepWs step_layout(int B, int hw, int D, int C, int H, int W) {
    StepWs w;
    const size_t M = (size_t)B * hw, Cp = class_pad(C);
    size_t o = 0;
    w.v_in = o;     o += al(M * D * 2);
    w.labels = o;   o += al((size_t)B * H * W * 8);
    w.packed = o;   o += al((size_t)B * H * W * 2);
    w.t_in = o;     o += al((size_t)C * D * 4);
    w.t_hat = o;    o += al(Cp * D * 2);
    w.inv_t = o;    o += al((size_t)C * 4);
    w.v_hat = o;    o += al(M * D * 2);
    w.inv_v = o;    o += al(M * 4);
    w.logits = o;   o += al(M * C * 4);
    w.grad_low = o; o += al(M * C * 4);
    w.grad_v = o;   o += al(M * D * 2);
    w.grad_t = o;   o += al((size_t)C * D * 4);
    w.bwd_ws = o;   o += al((size_t)lc2is_cosine_logits_bwd_workspace(B, hw, D, 1, C));
    w.confmat = o;  o += al((size_t)C * C * 8);
    w.scalars = o;  o += 256;     // [0] double loss_sum, [8] int64 n_valid, [16] float gscale, [20] float loss
    w.total = o;
    return w;
}
}  // namespace

extern "C" int64_t lc2is_head_step_workspace(int B, int hw, int D, int C, int H, int W) {
    return (int64_t)step_layout(B, hw, D, C, H, W).total;
}

extern "C" int lc2is_head_step_host(const void* h_v, const float* h_t, const int64_t* h_labels,
                                    int B, int h, int w, int D, int C, int H, int W,
                                    int64_t ignore_index, float logit_scale, int do_backward,
                                    float* h_out_loss, int64_t* h_out_n_valid, int64_t* h_out_confmat,
                                    void* d_ws, lc2is_stream_t stream, lc2is_stream_t copy_stream) {
    if (int e = ensure_device()) return e;
    if (!h_v || !h_t || !h_labels || !h_out_loss || !h_out_n_valid || !h_out_confmat || !d_ws)
        return fail(LC2IS_ERR_ARG, "null pointer%s");
    if (B <= 0) return fail(LC2IS_ERR_SHAPE, "B must be positive%s");
    cudaStream_t st = (cudaStream_t)stream;
    cudaStream_t cst = copy_stream ? (cudaStream_t)copy_stream : st;
    const bool piped = copy_stream && copy_stream != stream;
    const int hw = h * w;
    const StepWs L = step_layout(B, hw, D, C, H, W);
    uint8_t* ws = (uint8_t*)d_ws;
    const size_t M = (size_t)B * hw;
    uint8_t* d_v = ws + L.v_in;
    int64_t* d_labels = (int64_t*)(ws + L.labels);
    float* d_t = (float*)(ws + L.t_in);
    void* d_that = ws + L.t_hat;
    float* d_invt = (float*)(ws + L.inv_t);
    uint8_t* d_vhat = ws + L.v_hat;
    float* d_invv = (float*)(ws + L.inv_v);
    float* d_logits = (float*)(ws + L.logits);
    float* d_glow = (float*)(ws + L.grad_low);
    uint16_t* d_packed = (uint16_t*)(ws + L.packed);
    double* d_loss_sum = (double*)(ws + L.scalars);
    int64_t* d_nvalid = (int64_t*)(ws + L.scalars + 8);
    float* d_gscale = (float*)(ws + L.scalars + 16);
    float* d_loss = (float*)(ws + L.scalars + 20);
    int64_t* d_cm = (int64_t*)(ws + L.confmat);
    // Power-of-two scales 8 / 16 run the split form of K2 (label prepass + packed-label kernels); other
    // geometries the one-call K2 and the int64-label K3.
    const bool split = lc2is_ce_split_supported(h, w, H, W) != 0;

    // The batch is cut into chunks: chunk i+1 is copied host->device on `copy_stream` while the
    // kernels of chunk i run on `stream` (engine.py:75 / :145 copy the whole batch up front).
    // Gradients are produced un-normalised (g = 1) per chunk; the 1/N_valid of the 'mean' reduction is
    // only known after the last chunk and is applied by K1b through its device-side grad_scale.
    const int nchunk = piped ? (B < 4 ? B : 4) : 1;
    const int bc = (B + nchunk - 1) / nchunk;
    cudaEvent_t ev_start = nullptr, ev_copy[4] = {nullptr, nullptr, nullptr, nullptr};
    auto cleanup = [&]() {
        if (ev_start) cudaEventDestroy(ev_start);
        for (auto& e : ev_copy) if (e) cudaEventDestroy(e);
    };
#define STEP_CUDA(call)                                                  \
    do {                                                                 \
        cudaError_t e__ = (call);                                        \
        if (e__ != cudaSuccess) { cleanup(); return lc2is::cuda_fail(e__, #call); } \
    } while (0)
#define STEP_RC(call)                                                    \
    do {                                                                 \
        int rc__ = (call);                                               \
        if (rc__) { cleanup(); return rc__; }                            \
    } while (0)
    if (piped) {
        STEP_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        for (int i = 0; i < nchunk; ++i) STEP_CUDA(cudaEventCreateWithFlags(&ev_copy[i], cudaEventDisableTiming));
        STEP_CUDA(cudaEventRecord(ev_start, st));           // the workspace is free once earlier work on
        STEP_CUDA(cudaStreamWaitEvent(cst, ev_start, 0));   // `stream` has drained
    }
    STEP_CUDA(cudaMemsetAsync(ws + L.scalars, 0, 256, st));
    STEP_CUDA(cudaMemsetAsync(d_cm, 0, (size_t)C * C * 8, st));
    if (split && do_backward) STEP_CUDA(cudaMemsetAsync(d_glow, 0, M * C * 4, st));
    STEP_CUDA(cudaMemcpyAsync(d_t, h_t, (size_t)C * D * 4, cudaMemcpyHostToDevice, cst));
    for (int i = 0; i < nchunk; ++i) {
        const int b0 = i * bc, nb = (b0 + bc <= B ? bc : B - b0);
        if (nb <= 0) break;
        const size_t lab_off = (size_t)b0 * H * W, v_off = (size_t)b0 * hw * D * 2;
        STEP_CUDA(cudaMemcpyAsync(d_labels + lab_off, h_labels + lab_off, (size_t)nb * H * W * 8,
                                  cudaMemcpyHostToDevice, cst));
        STEP_CUDA(cudaMemcpyAsync(d_v + v_off, (const uint8_t*)h_v + v_off, (size_t)nb * hw * D * 2,
                                  cudaMemcpyHostToDevice, cst));
        if (piped) {
            STEP_CUDA(cudaEventRecord(ev_copy[i], cst));
            STEP_CUDA(cudaStreamWaitEvent(st, ev_copy[i], 0));
        }
        float* lg = d_logits + (size_t)b0 * C * hw;
        float* gl = do_backward ? d_glow + (size_t)b0 * C * hw : nullptr;
        if (split)
            STEP_RC(lc2is_ce_labels_prepass(d_labels + lab_off, nb, C, h, w, H, W, ignore_index, d_packed + lab_off,
                                            d_nvalid, gl, stream));
        else
            STEP_RC(lc2is_count_valid(d_labels + lab_off, (int64_t)nb * H * W, ignore_index, d_nvalid, stream));
        if (i == 0) STEP_RC(lc2is_proto_normalize(d_t, 1, C, D, 1, d_that, d_invt, stream));
        STEP_RC(lc2is_cosine_logits_fwd(d_v + v_off, LC2IS_BF16, nb, hw, D, d_that, 1, C, 1, logit_scale,
                                        d_vhat + v_off, d_invv + (size_t)b0 * hw, lg, stream));
        if (split) {
            STEP_RC(lc2is_upsample_ce_packed(lg, d_packed + lab_off, nb, C, h, w, H, W, d_loss_sum, gl, stream));
            STEP_RC(lc2is_argmax_confmat_lowres_packed(lg, nb, C, h, w, H, W, d_packed + lab_off, d_cm, nullptr,
                                                       nullptr, stream));
        } else {
            STEP_RC(lc2is_upsample_ce_fwd_bwd(lg, d_labels + lab_off, nb, C, h, w, H, W, ignore_index, nullptr,
                                              d_loss_sum, gl, nullptr, stream));
            STEP_RC(lc2is_argmax_confmat_lowres(lg, nb, C, h, w, H, W, LC2IS_BILINEAR, d_labels + lab_off, H, W,
                                                d_cm, nullptr, nullptr, stream));
        }
    }
    STEP_RC(lc2is_mean_scale(d_nvalid, 1.0f, d_gscale, stream));
    if (do_backward) {
        STEP_CUDA(cudaMemsetAsync(ws + L.grad_t, 0, (size_t)C * D * 4, st));
        STEP_RC(lc2is_cosine_logits_bwd(d_glow, LC2IS_F32, d_logits, d_vhat, d_invv, d_that, d_invt, B, hw, D, 1, C,
                                        1, logit_scale, d_gscale, ws + L.grad_v, LC2IS_BF16,
                                        (float*)(ws + L.grad_t), ws + L.bwd_ws, stream));
    }
    STEP_RC(lc2is_finalize_loss(d_loss_sum, d_nvalid, d_loss, stream));
    // D2H of the step's results (engine.py:108 .item(); :162-163)
    STEP_CUDA(cudaMemcpyAsync(h_out_loss, d_loss, 4, cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaMemcpyAsync(h_out_n_valid, d_nvalid, 8, cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaMemcpyAsync(h_out_confmat, d_cm, (size_t)C * C * 8, cudaMemcpyDeviceToHost, st));
    STEP_CUDA(cudaStreamSynchronize(st));
    cleanup();
#undef STEP_CUDA
#undef STEP_RC
    return 0;
}
