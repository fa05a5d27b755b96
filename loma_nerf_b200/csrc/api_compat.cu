// api_compat.cu -- the five ctypes-facing symbols of the loma-compiled _code/nerf.so and
// _code/mlp_fit.so (include/loma_nerf_b200.h section 1), as gather -> flat host step -> scatter.
//
// Buffer conventions followed (/root/reference/mlp_utils.py:33-118 builds them,
// /root/reference/loma_public/compiler.py:262-276 types them): float** / float*** are tables of
// independently allocated rows; the callee owns nothing; forward scratch arrays are written in
// place; every d_ buffer is accumulated into; int adjoints are never touched; the grad functions
// leave the primal scratch arrays exactly as they were on entry.
//
// Shapes: dims[0] = layer_input_w, dims[l+1] = weight_shapes[l][1]; rows the bias/activation loops
// run over = intermediate_output_shapes[0][0] (the reference hosts pass the same row count for
// every layer, train_nerf.py:230-234).  Errors never abort: the loss comes back NaN and the grad
// functions NaN-fill d_ws so the hosts' NaN guard (train_nerf.py:486-489) trips.
#include <math.h>
#include <string.h>

#include <mutex>
#include <vector>

#include "lnb_internal.h"

namespace {

std::mutex g_compat_mu;

struct Shapes {
    lnb_mlp mlp;
    int N, R, S, M, Wt;
    bool ok;
};

Shapes read_shapes(int layer_input_h, int layer_input_w, int target_h, int target_w, int num_weights,
                   int **weight_shapes, int **inter_shapes, int num_samples, int head)
{
    Shapes s{};
    s.ok = false;
    if (num_weights < 1 || num_weights > LNB_MAX_LAYERS || !weight_shapes || !inter_shapes) return s;
    if (layer_input_h < 0 || layer_input_w < 1 || target_h < 0 || target_w < 1 || num_samples < 1) return s;
    s.mlp.n_layers = num_weights;
    s.mlp.head = head;
    s.mlp.dims[0] = layer_input_w;
    s.mlp.max_in = layer_input_w;
    s.mlp.max_out = 1;
    for (int l = 0; l < num_weights; ++l) {
        int out_l = weight_shapes[l][1];
        if (out_l < 1) return s;
        s.mlp.dims[l + 1] = out_l;
        if (out_l > s.mlp.max_out) s.mlp.max_out = out_l;
        if (l + 1 < num_weights && out_l > s.mlp.max_in) s.mlp.max_in = out_l;
    }
    s.N = layer_input_h;
    s.R = target_h;
    s.S = num_samples;
    s.Wt = target_w;
    s.M = inter_shapes[0][0] > s.N ? inter_shapes[0][0] : s.N;
    s.ok = true;
    return s;
}

void gather2(std::vector<float> &dst, float **rows, size_t n_rows, size_t n_cols, size_t ld)
{
    for (size_t i = 0; i < n_rows; ++i) memcpy(&dst[i * ld], rows[i], n_cols * sizeof(float));
}

// returns true when any gathered value is non-zero
bool gather_inter(std::vector<float> &dst, float ***inter, const Shapes &s)
{
    bool nz = false;
    const size_t ld = s.mlp.max_out;
    for (int l = 0; l < s.mlp.n_layers; ++l)
        for (int i = 0; i < s.M; ++i) {
            const float *row = inter[l][i];
            float *o = &dst[((size_t)l * s.M + i) * ld];
            for (int j = 0; j < s.mlp.dims[l + 1]; ++j) {
                o[j] = row[j];
                nz |= (row[j] != 0.0f);
            }
        }
    return nz;
}

void gather_weights(std::vector<float> &w, std::vector<float> &b, float ***ws, float **bs, const Shapes &s)
{
    const lnb_mlp &m = s.mlp;
    for (int l = 0; l < m.n_layers; ++l) {
        for (int k = 0; k < m.dims[l]; ++k)
            memcpy(&w[((size_t)l * m.max_in + k) * m.max_out], ws[l][k], m.dims[l + 1] * sizeof(float));
        memcpy(&b[(size_t)l * m.max_out], bs[l], m.dims[l + 1] * sizeof(float));
    }
}

struct Flat {
    std::vector<float> X, w, b, target, dists, inter, color;
    std::vector<float> rgba, alpha, cumprod, weights;
    std::vector<float> d_w, d_b, d_X, d_target, d_dists, d_color, d_inter;
    float loss = NAN;
};

void nan_fill_d_ws(float ***d_ws, const Shapes &s)
{
    if (!d_ws || !s.ok) return;
    for (int l = 0; l < s.mlp.n_layers; ++l)
        for (int k = 0; k < s.mlp.dims[l]; ++k)
            for (int j = 0; j < s.mlp.dims[l + 1]; ++j) d_ws[l][k][j] = NAN;
}

// common: gather inputs, run, leave results in `f`.  grad = also the backward pass.
int run(const Shapes &s, bool nerf, bool grad, float seed, float **layer_input, float ***ws,
        float **bs, float **target, float **dists, float ***inter, float **color_in, Flat &f,
        bool want_fwd_scratch)
{
    lnb_ctx *ctx = lnb_default_ctx();
    if (!ctx) return LNB_ERR_CUDA;
    const lnb_mlp &m = s.mlp;
    const size_t N = s.N, R = s.R, S = nerf ? s.S : 1, M = s.M, L = m.n_layers;
    f.X.assign(N * m.dims[0], 0.f);
    gather2(f.X, layer_input, N, m.dims[0], m.dims[0]);
    f.w.assign(L * (size_t)m.max_in * m.max_out, 0.f);
    f.b.assign(L * (size_t)m.max_out, 0.f);
    gather_weights(f.w, f.b, ws, bs, s);
    f.target.assign(R * s.Wt, 0.f);
    gather2(f.target, target, R, s.Wt, s.Wt);
    lnb_step_args a{};
    a.R = s.R; a.S = (int)S; a.n_rows = s.N; a.rows = s.M; a.target_w = s.Wt;
    a.X = f.X.data(); a.ws = f.w.data(); a.bs = f.b.data(); a.target = f.target.data();
    a.inter_rows = s.M; a.inter_ld = m.max_out;
    f.inter.assign(L * M * m.max_out, 0.f);
    bool inter_nz = inter ? gather_inter(f.inter, inter, s) : false;
    if (want_fwd_scratch || inter_nz) {
        a.inter = f.inter.data();
        a.inter_accumulate = inter_nz ? 1 : 0;
    }
    if (nerf) {
        f.dists.assign(R * S, 0.f);
        gather2(f.dists, dists, R, S, S);
        a.dists = f.dists.data();
        f.color.assign(R * 3, 0.f);
        if (color_in) gather2(f.color, color_in, R, 3, 3);
        a.color = f.color.data();
        a.color_accumulate = 1;
        if (want_fwd_scratch) {
            f.rgba.assign(R * S * 4, 0.f); f.alpha.assign(R * S, 0.f);
            f.cumprod.assign(R * S, 0.f); f.weights.assign(R * S, 0.f);
            a.rgba = f.rgba.data(); a.alpha = f.alpha.data();
            a.cumprod = f.cumprod.data(); a.weights = f.weights.data();
        }
    }
    a.loss = &f.loss;
    a.path = LNB_PATH_F32_LAYERWISE;
    if (grad) {
        a.want_grad = 1; a.seed_mode = LNB_SEED_VALUE; a.seed = seed;
        f.d_w.assign(f.w.size(), 0.f); f.d_b.assign(f.b.size(), 0.f);
        f.d_X.assign(f.X.size(), 0.f); f.d_target.assign(f.target.size(), 0.f);
        f.d_color.assign(R * s.Wt, 0.f); f.d_inter.assign(f.inter.size(), 0.f);
        a.d_ws = f.d_w.data(); a.d_bs = f.d_b.data(); a.d_X = f.d_X.data();
        a.d_target = f.d_target.data(); a.d_color = f.d_color.data(); a.d_inter = f.d_inter.data();
        if (nerf) { f.d_dists.assign(R * S, 0.f); a.d_dists = f.d_dists.data(); }
    }
    return nerf ? lnb_nerf_step_host(ctx, &m, &a) : lnb_fit_step_host(ctx, &m, &a);
}

void scatter_inter(const std::vector<float> &src, float ***inter, const Shapes &s, int n_rows, bool add)
{
    const size_t ld = s.mlp.max_out;
    for (int l = 0; l < s.mlp.n_layers; ++l)
        for (int i = 0; i < n_rows; ++i) {
            float *row = inter[l][i];
            const float *v = &src[((size_t)l * s.M + i) * ld];
            for (int j = 0; j < s.mlp.dims[l + 1]; ++j) row[j] = add ? row[j] + v[j] : v[j];
        }
}

void scatter_add2(const std::vector<float> &src, float **rows, size_t n_rows, size_t n_cols)
{
    if (!rows) return;
    for (size_t i = 0; i < n_rows; ++i)
        for (size_t j = 0; j < n_cols; ++j) rows[i][j] += src[i * n_cols + j];
}

void scatter_grads(const Flat &f, const Shapes &s, bool nerf, float **d_layer_input, float ***d_ws,
                   float **d_bs, float **d_target, float **d_dists, float **d_color,
                   float ***d_inter)
{
    const lnb_mlp &m = s.mlp;
    scatter_add2(f.d_X, d_layer_input, s.N, m.dims[0]);
    for (int l = 0; l < m.n_layers; ++l) {
        if (d_ws)
            for (int k = 0; k < m.dims[l]; ++k)
                for (int j = 0; j < m.dims[l + 1]; ++j)
                    d_ws[l][k][j] += f.d_w[((size_t)l * m.max_in + k) * m.max_out + j];
        if (d_bs)
            for (int j = 0; j < m.dims[l + 1]; ++j) d_bs[l][j] += f.d_b[(size_t)l * m.max_out + j];
    }
    scatter_add2(f.d_target, d_target, s.R, s.Wt);
    if (nerf) {
        scatter_add2(f.d_dists, d_dists, s.R, s.S);
        scatter_add2(f.d_color, d_color, s.R, 3);
    }
    if (d_inter) scatter_inter(f.d_inter, d_inter, s, nerf ? s.R * s.S : s.R, true);
}

} // namespace

extern "C" float nerf_evaluate_and_march(float **layer_input, int layer_input_h, int layer_input_w,
                                         float ***ws, float **bs, float **target_image,
                                         int target_image_h, int target_image_w, int num_weights,
                                         int **weight_shapes, int **bias_shapes,
                                         int **intermediate_output_shapes,
                                         float ***intermediate_outputs,
                                         float ***img_sample_rgba_arr, int num_samples,
                                         float **dists, float **alpha, float **cumprod_alpha,
                                         float **weights_samples, float **accumulated_color)
{
    (void)bias_shapes;
    std::lock_guard<std::mutex> lk(g_compat_mu);
    Shapes s = read_shapes(layer_input_h, layer_input_w, target_image_h, target_image_w, num_weights,
                           weight_shapes, intermediate_output_shapes, num_samples, LNB_HEAD_NERF);
    if (!s.ok || target_image_w != 3 || (long long)s.R * s.S > s.N) return NAN;
    Flat f;
    if (run(s, true, false, 0.f, layer_input, ws, bs, target_image, dists, intermediate_outputs,
            accumulated_color, f, true) != LNB_OK)
        return NAN;
    scatter_inter(f.inter, intermediate_outputs, s, s.M, false);
    for (int r = 0; r < s.R; ++r) {
        for (int sm = 0; sm < s.S; ++sm) {
            size_t i = (size_t)r * s.S + sm;
            memcpy(img_sample_rgba_arr[r][sm], &f.rgba[i * 4], 4 * sizeof(float));
            alpha[r][sm] = f.alpha[i];
            cumprod_alpha[r][sm] = f.cumprod[i];
            weights_samples[r][sm] = f.weights[i];
        }
        memcpy(accumulated_color[r], &f.color[(size_t)r * 3], 3 * sizeof(float));
    }
    return f.loss;
}

extern "C" void grad_nerf_evaluate_and_march(
    float **layer_input, float **d_layer_input, int layer_input_h, int *d_layer_input_h,
    int layer_input_w, int *d_layer_input_w, float ***ws, float ***d_ws, float **bs, float **d_bs,
    float **target_image, float **d_target_image, int target_image_h, int *d_target_image_h,
    int target_image_w, int *d_target_image_w, int num_weights, int *d_num_weights,
    int **weight_shapes, int **d_weight_shapes, int **bias_shapes, int **d_bias_shapes,
    int **intermediate_output_shapes, int **d_intermediate_output_shapes,
    float ***intermediate_outputs, float ***d_intermediate_outputs, float ***img_sample_rgba_arr,
    float ***d_img_sample_rgba_arr, int num_samples, int *d_num_samples, float **dists,
    float **d_dists, float **alpha, float **d_alpha, float **cumprod_alpha,
    float **d_cumprod_alpha, float **weights_samples, float **d_weights_samples,
    float **accumulated_color, float **d_accumulated_color, float _dreturn)
{
    (void)d_layer_input_h; (void)d_layer_input_w; (void)d_target_image_h; (void)d_target_image_w;
    (void)d_num_weights; (void)d_weight_shapes; (void)bias_shapes; (void)d_bias_shapes;
    (void)d_intermediate_output_shapes; (void)img_sample_rgba_arr; (void)d_img_sample_rgba_arr;
    (void)d_num_samples; (void)alpha; (void)d_alpha; (void)cumprod_alpha; (void)d_cumprod_alpha;
    (void)weights_samples; (void)d_weights_samples;
    std::lock_guard<std::mutex> lk(g_compat_mu);
    Shapes s = read_shapes(layer_input_h, layer_input_w, target_image_h, target_image_w, num_weights,
                           weight_shapes, intermediate_output_shapes, num_samples, LNB_HEAD_NERF);
    if (!s.ok || target_image_w != 3 || (long long)s.R * s.S > s.N) { nan_fill_d_ws(d_ws, s); return; }
    Flat f;
    if (run(s, true, true, _dreturn, layer_input, ws, bs, target_image, dists, intermediate_outputs,
            accumulated_color, f, false) != LNB_OK) {
        nan_fill_d_ws(d_ws, s);
        return;
    }
    scatter_grads(f, s, true, d_layer_input, d_ws, d_bs, d_target_image, d_dists,
                  d_accumulated_color, d_intermediate_outputs);
}

extern "C" float mlp_fit(float **layer_input, int layer_input_h, int layer_input_w,
                         float **layer_output, float ***ws, float **bs, float **target_image,
                         int target_image_h, int target_image_w, int num_weights,
                         int **weight_shapes, int **bias_shapes, int **intermediate_output_shapes,
                         float ***intermediate_outputs)
{
    (void)layer_output; (void)bias_shapes;
    std::lock_guard<std::mutex> lk(g_compat_mu);
    Shapes s = read_shapes(layer_input_h, layer_input_w, target_image_h, target_image_w, num_weights,
                           weight_shapes, intermediate_output_shapes, 1, LNB_HEAD_SIGMOID);
    if (!s.ok || s.R > s.N || s.Wt > s.mlp.dims[s.mlp.n_layers]) return NAN;
    Flat f;
    if (run(s, false, false, 0.f, layer_input, ws, bs, target_image, nullptr, intermediate_outputs,
            nullptr, f, true) != LNB_OK)
        return NAN;
    scatter_inter(f.inter, intermediate_outputs, s, s.M, false);
    return f.loss;
}

extern "C" void grad_mlp_fit(float **layer_input, float **d_layer_input, int layer_input_h,
                             int *d_layer_input_h, int layer_input_w, int *d_layer_input_w,
                             float **layer_output, float **d_layer_output, float ***ws,
                             float ***d_ws, float **bs, float **d_bs, float **target_image,
                             float **d_target_image, int target_image_h, int *d_target_image_h,
                             int target_image_w, int *d_target_image_w, int num_weights,
                             int *d_num_weights, int **weight_shapes, int **d_weight_shapes,
                             int **bias_shapes, int **d_bias_shapes,
                             int **intermediate_output_shapes, int **d_intermediate_output_shapes,
                             float ***intermediate_outputs, float ***d_intermediate_outputs,
                             float _dreturn)
{
    (void)d_layer_input_h; (void)d_layer_input_w; (void)layer_output; (void)d_layer_output;
    (void)d_target_image_h; (void)d_target_image_w; (void)d_num_weights; (void)d_weight_shapes;
    (void)bias_shapes; (void)d_bias_shapes; (void)d_intermediate_output_shapes;
    std::lock_guard<std::mutex> lk(g_compat_mu);
    Shapes s = read_shapes(layer_input_h, layer_input_w, target_image_h, target_image_w, num_weights,
                           weight_shapes, intermediate_output_shapes, 1, LNB_HEAD_SIGMOID);
    if (!s.ok || s.R > s.N || s.Wt > s.mlp.dims[s.mlp.n_layers]) { nan_fill_d_ws(d_ws, s); return; }
    Flat f;
    if (run(s, false, true, _dreturn, layer_input, ws, bs, target_image, nullptr,
            intermediate_outputs, nullptr, f, false) != LNB_OK) {
        nan_fill_d_ws(d_ws, s);
        return;
    }
    scatter_grads(f, s, false, d_layer_input, d_ws, d_bs, d_target_image, nullptr, nullptr,
                  d_intermediate_outputs);
}

extern "C" void mult_a_b(float **a, int a_h, int a_w, float **b, int b_h, int b_w, float **c)
{
    std::lock_guard<std::mutex> lk(g_compat_mu);
    lnb_ctx *ctx = lnb_default_ctx();
    auto fail = [&]() {
        for (int i = 0; i < a_h; ++i)
            for (int j = 0; j < b_w; ++j) c[i][j] = NAN;
    };
    if (!ctx || a_h < 0 || a_w < 0 || b_w < 0 || b_h < a_w) { if (a_h > 0 && b_w > 0) fail(); return; }
    if (a_h == 0 || b_w == 0) return;
    const size_t na = (size_t)a_h * a_w, nb = (size_t)a_w * b_w, nc = (size_t)a_h * b_w;
    std::vector<float> h(na + nb + nc);
    for (int i = 0; i < a_h; ++i) memcpy(&h[(size_t)i * a_w], a[i], a_w * sizeof(float));
    for (int k = 0; k < a_w; ++k) memcpy(&h[na + (size_t)k * b_w], b[k], b_w * sizeof(float));
    for (int i = 0; i < a_h; ++i) memcpy(&h[na + nb + (size_t)i * b_w], c[i], b_w * sizeof(float));
    float *dev = nullptr;
    bool ok = cudaSetDevice(ctx->device) == cudaSuccess &&
              cudaMalloc((void **)&dev, h.size() * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(dev, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    ok = ok && lnb_mult_a_b(ctx, dev, a_h, a_w, dev + na, b_w, dev + na + nb) == LNB_OK;
    ok = ok && cudaMemcpyAsync(&h[na + nb], dev + na + nb, nc * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess;
    ok = ok && cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    if (dev) cudaFree(dev);
    if (!ok) { cudaGetLastError(); fail(); return; }
    for (int i = 0; i < a_h; ++i) memcpy(c[i], &h[na + nb + (size_t)i * b_w], b_w * sizeof(float));
}

// autodiff.py:199-208: value / tangent pair constructor present in every loma-generated library
extern "C" LNB_API _dfloat make__dfloat(float val, float dval)
{
    _dfloat r;
    r.val = val;
    r.dval = dval;
    return r;
}
