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
    size_t v_in, labels, t_in, t_hat, inv_t, v_hat, inv_v, logits, grad_low, grad_bf16, grad_v, grad_t,
        bwd_ws, confmat, scalars, total;
};
inline size_t al(size_t x) { return (x + 255) / 256 * 256; }
StepWs step_layout(int B, int hw, int D, int C, int H, int W) {
    StepWs w;
    const size_t M = (size_t)B * hw, Cp = class_pad(C);
    size_t o = 0;
    w.v_in = o;     o += al(M * D * 2);
    w.labels = o;   o += al((size_t)B * H * W * 8);
    w.t_in = o;     o += al((size_t)C * D * 4);
    w.t_hat = o;    o += al(Cp * D * 2);
    w.inv_t = o;    o += al((size_t)C * 4);
    w.v_hat = o;    o += al(M * D * 2);
    w.inv_v = o;    o += al(M * 4);
    w.logits = o;   o += al(M * C * 4);
    w.grad_low = o; o += al(M * C * 4);
    w.grad_bf16 = o; o += al(M * Cp * 2);
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
                                    void* d_ws, lc2is_stream_t stream) {
    if (int e = ensure_device()) return e;
    if (!h_v || !h_t || !h_labels || !h_out_loss || !h_out_n_valid || !h_out_confmat || !d_ws)
        return fail(LC2IS_ERR_ARG, "null pointer%s");
    cudaStream_t st = (cudaStream_t)stream;
    const int hw = h * w;
    const StepWs L = step_layout(B, hw, D, C, H, W);
    uint8_t* ws = (uint8_t*)d_ws;
    const size_t M = (size_t)B * hw;
    void* d_v = ws + L.v_in;
    int64_t* d_labels = (int64_t*)(ws + L.labels);
    float* d_t = (float*)(ws + L.t_in);
    void* d_that = ws + L.t_hat;
    float* d_invt = (float*)(ws + L.inv_t);
    void* d_vhat = ws + L.v_hat;
    float* d_invv = (float*)(ws + L.inv_v);
    float* d_logits = (float*)(ws + L.logits);
    float* d_glow = (float*)(ws + L.grad_low);
    void* d_gbf = ws + L.grad_bf16;
    double* d_loss_sum = (double*)(ws + L.scalars);
    int64_t* d_nvalid = (int64_t*)(ws + L.scalars + 8);
    float* d_gscale = (float*)(ws + L.scalars + 16);
    float* d_loss = (float*)(ws + L.scalars + 20);
    int64_t* d_cm = (int64_t*)(ws + L.confmat);

    // H2D of the batch (engine.py:75 / :145)
    LC2IS_CUDA(cudaMemcpyAsync(d_labels, h_labels, (size_t)B * H * W * 8, cudaMemcpyHostToDevice, st));
    LC2IS_CUDA(cudaMemcpyAsync(d_t, h_t, (size_t)C * D * 4, cudaMemcpyHostToDevice, st));
    LC2IS_CUDA(cudaMemcpyAsync(d_v, h_v, M * D * 2, cudaMemcpyHostToDevice, st));
    LC2IS_CUDA(cudaMemsetAsync(ws + L.scalars, 0, 256, st));
    LC2IS_CUDA(cudaMemsetAsync(d_cm, 0, (size_t)C * C * 8, st));

    if (int e = lc2is_count_valid(d_labels, (int64_t)B * H * W, ignore_index, d_nvalid, stream)) return e;
    if (int e = lc2is_mean_scale(d_nvalid, 1.0f, d_gscale, stream)) return e;
    if (int e = lc2is_proto_normalize(d_t, 1, C, D, 1, d_that, d_invt, stream)) return e;
    if (int e = lc2is_cosine_logits_fwd(d_v, LC2IS_BF16, B, hw, D, d_that, 1, C, 1, logit_scale, d_vhat, d_invv,
                                        d_logits, stream)) return e;
    if (int e = lc2is_upsample_ce_fwd_bwd(d_logits, d_labels, B, C, h, w, H, W, ignore_index, d_gscale, d_loss_sum,
                                          d_glow, do_backward ? d_gbf : nullptr, stream)) return e;
    if (do_backward) {
        LC2IS_CUDA(cudaMemsetAsync(ws + L.grad_t, 0, (size_t)C * D * 4, st));
        if (int e = lc2is_cosine_logits_bwd(d_gbf, d_logits, d_vhat, d_invv, d_that, d_invt, B, hw, D, 1, C, 1,
                                            logit_scale, nullptr, ws + L.grad_v, LC2IS_BF16,
                                            (float*)(ws + L.grad_t), ws + L.bwd_ws, stream)) return e;
    }
    if (int e = lc2is_argmax_confmat_lowres(d_logits, B, C, h, w, H, W, LC2IS_BILINEAR, d_labels, H, W, d_cm,
                                            nullptr, nullptr, stream)) return e;
    if (int e = lc2is_finalize_loss(d_loss_sum, d_nvalid, d_loss, stream)) return e;
    // D2H of the step's results (engine.py:108 .item(); :162-163)
    LC2IS_CUDA(cudaMemcpyAsync(h_out_loss, d_loss, 4, cudaMemcpyDeviceToHost, st));
    LC2IS_CUDA(cudaMemcpyAsync(h_out_n_valid, d_nvalid, 8, cudaMemcpyDeviceToHost, st));
    LC2IS_CUDA(cudaMemcpyAsync(h_out_confmat, d_cm, (size_t)C * C * 8, cudaMemcpyDeviceToHost, st));
    LC2IS_CUDA(cudaStreamSynchronize(st));
    return 0;
}
