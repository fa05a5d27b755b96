/*
 * oracle/nerf_oracle.c -- CPU restatement of the loma-nerf hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA
 * path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may compile, link or call it.  The product library
 * (loma_nerf_b200/csrc) never includes or calls anything in oracle/.
 *
 * Parity pin: this restatement is checked (tests/test_oracle.py) against
 *   (1) the reference's only known-answer test on this path, mult_a_b ->
 *       [[500],[1100],[1700]]            (/root/reference/fit_img.py:363-374),
 *   (2) the PE identity-prefix asserts   (/root/reference/pos_encoding.py:34,68),
 *   (3) outputs of the REAL reference (its loma programs compiled by its own
 *       compiler, oracle/build_ref.py -> oracle/_ref/) recorded as golden
 *       vectors in the tests/golden/ npz files by tests/golden/make_golden.py, and
 *       live against the .so files in oracle/_ref/ whenever those files are present.
 *
 * Everything is IEEE fp32 with the reference's operation order (sequential k,
 * separate multiply and add: build with -ffp-contract=off), int32 indices.
 * Buffers are FLAT row-major here; the reference's ragged float** rows hold the
 * same values (mlp_utils.py:33-118).
 *
 * Layouts:  X [N][C_in];  ws [L][max_in][max_out] (row = input channel,
 * mlp_utils.py:166-204,272-313);  bs [L][max_out];  dims[l] = in_l,
 * dims[l+1] = out_l;  inter [L][rows][ld] with rows >= N (the reference host
 * passes rows = 256 and a 256-wide scratch, train_nerf.py:230-238);
 * rgba [R][S][4]; dists/alpha/cumprod/weights [R][S]; acc, target [R][3].
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define W_AT(ws, l, k, j) (ws)[((size_t)(l) * max_in + (k)) * max_out + (j)]
#define B_AT(bs, l, j) (bs)[(size_t)(l) * max_out + (j)]
#define I_AT(inter, l, i, j) (inter)[((size_t)(l) * rows + (i)) * ld + (j)]

/* head kinds */
#define HEAD_NERF 0    /* sigmoid on channels != 3, ReLU on channel 3 (scripts/nerf.py:147-167) */
#define HEAD_SIGMOID 1 /* sigmoid on every channel             (scripts/mlp_fit.py:121-132) */

/* ---- MLP forward: scripts/nerf.py:67-167 == scripts/mlp_fit.py:39-132 ---------------------- */
static void mlp_forward(const float *X, int N, int C_in, const float *ws, const float *bs, int L,
                        const int *dims, int max_in, int max_out, int rows, int ld, float *inter,
                        int head)
{
    for (int l = 0; l < L; ++l) {
        int in_l = dims[l], out_l = dims[l + 1];
        if (l == 0) {
            /* nerf.py:81-89: rows < layer_input_h, k < layer_input_w */
            for (int i = 0; i < N; ++i)
                for (int j = 0; j < out_l; ++j)
                    for (int k = 0; k < C_in; ++k) {
                        float p = X[(size_t)i * C_in + k] * W_AT(ws, 0, k, j);
                        I_AT(inter, 0, i, j) = I_AT(inter, 0, i, j) + p;
                    }
        } else {
            /* nerf.py:108-116: rows < intermediate_output_shapes[l-1][0] */
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < out_l; ++j)
                    for (int k = 0; k < in_l; ++k) {
                        float p = I_AT(inter, l - 1, i, k) * W_AT(ws, l, k, j);
                        I_AT(inter, l, i, j) = I_AT(inter, l, i, j) + p;
                    }
        }
        /* bias, nerf.py:95-100 / 122-127: rows < intermediate_output_shapes[l][0] */
        for (int i = 0; i < rows; ++i)
            for (int j = 0; j < out_l; ++j)
                I_AT(inter, l, i, j) = I_AT(inter, l, i, j) + B_AT(bs, l, j);
        if (l < L - 1) {
            /* ReLU in place, nerf.py:134-146 */
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < out_l; ++j)
                    if (!(I_AT(inter, l, i, j) > 0.0f)) I_AT(inter, l, i, j) = 0.0f;
        } else {
            for (int i = 0; i < rows; ++i)
                for (int j = 0; j < out_l; ++j) {
                    float z = I_AT(inter, l, i, j);
                    if (head == HEAD_NERF && j == 3) {
                        if (!(z > 0.0f)) z = 0.0f; /* nerf.py:157-162 */
                    } else {
                        z = 1.0f / (1.0f + expf(0.0f - z)); /* nerf.py:165, mlp_fit.py:130 */
                    }
                    I_AT(inter, l, i, j) = z;
                }
        }
    }
}

/* ---- compositing forward: scripts/nerf.py:176-288 ------------------------------------------ */
static void composite_forward(const float *head_out, int ld, int R, int S, const float *dists,
                              float *rgba, float *alpha, float *cumprod, float *weights, float *acc)
{
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s)
            for (int c = 0; c < 4; ++c) /* nerf.py:182-191 */
                rgba[((size_t)r * S + s) * 4 + c] = head_out[((size_t)r * S + s) * ld + c];
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) { /* nerf.py:200-205 */
            size_t i = (size_t)r * S + s;
            alpha[i] = 1.0f - expf((0.0f - rgba[i * 4 + 3]) * dists[i]);
        }
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) { /* nerf.py:215-220 */
            size_t i = (size_t)r * S + s;
            cumprod[i] = (1.0f - alpha[i]) + 1e-10f;
        }
    for (int r = 0; r < R; ++r)
        for (int s = 1; s < S; ++s) { /* inclusive product, nerf.py:226-232 */
            size_t i = (size_t)r * S + s;
            cumprod[i] = cumprod[i - 1] * cumprod[i];
        }
    for (int r = 0; r < R; ++r) /* nerf.py:252-258 (238-246 is a dead store) */
        cumprod[(size_t)r * S] = 1.0f;
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) { /* nerf.py:267-272 */
            size_t i = (size_t)r * S + s;
            weights[i] = alpha[i] * cumprod[i];
        }
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) { /* nerf.py:281-288, accumulates onto caller contents */
            size_t i = (size_t)r * S + s;
            for (int c = 0; c < 3; ++c)
                acc[r * 3 + c] = acc[r * 3 + c] + weights[i] * rgba[i * 4 + c];
        }
}

/* SSE loss: nerf.py:297-302, mlp_fit.py:140-145. pred rows have stride ldp. */
static float sse_loss(const float *pred, int ldp, const float *target, int R, int Wt)
{
    float loss = 0.0f;
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < Wt; ++c) {
            float d = pred[(size_t)r * ldp + c] - target[(size_t)r * Wt + c];
            loss = loss + d * d;
        }
    return loss;
}

float oracle_nerf_forward(const float *X, int N, int C_in, const float *ws, const float *bs, int L,
                          const int *dims, int max_in, int max_out, const float *target, int R,
                          int S, const float *dists, int rows, int ld, float *inter, float *rgba,
                          float *alpha, float *cumprod, float *weights, float *acc)
{
    mlp_forward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_NERF);
    composite_forward(&I_AT(inter, L - 1, 0, 0), ld, R, S, dists, rgba, alpha, cumprod, weights,
                      acc);
    return sse_loss(acc, 3, target, R, 3);
}

float oracle_mlp_fit_forward(const float *X, int N, int C_in, const float *ws, const float *bs,
                             int L, const int *dims, int max_in, int max_out, const float *target,
                             int R, int Wt, int rows, int ld, float *inter)
{
    mlp_forward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_SIGMOID);
    return sse_loss(&I_AT(inter, L - 1, 0, 0), ld, target, R, Wt);
}

/* ---- MLP backward: reversal of nerf.py:67-167 by reverse_diff.py:576-616,673-696,746-951.
 * `post` holds post-activation values (what the forward leaves in intermediate_outputs); since
 * relu(z) > 0 <=> z > 0 the ReLU masks are recovered from it; for sigmoid the loma rule
 * differentiates 1/(1+exp(0-z)) wrt the taped z: d z = d y * e/(1+e)^2 with e = expf(-z); we
 * recompute from y: e = 1/y - 1 is avoided; instead y*(1-y) (equal in exact arithmetic).
 * dZ enters in d_inter[L-1] rows (already the adjoint of the head's POST-activation values) and
 * every d_inter[l] leaves holding the pre-activation adjoint dZ_l. */
static void mlp_backward(const float *X, int N, int C_in, const float *ws, const float *bs, int L,
                         const int *dims, int max_in, int max_out, int rows, int ld,
                         const float *post, int head, float *d_inter, float *d_ws, float *d_bs,
                         float *d_X)
{
    (void)bs;
    for (int l = L - 1; l >= 0; --l) {
        int in_l = dims[l], out_l = dims[l + 1];
        /* activation reversal (in place on the adjoint) */
        for (int i = rows - 1; i >= 0; --i)
            for (int j = out_l - 1; j >= 0; --j) {
                float y = post[((size_t)l * rows + i) * ld + j];
                float d = I_AT(d_inter, l, i, j);
                if (l < L - 1 || (head == HEAD_NERF && j == 3)) {
                    if (!(y > 0.0f)) d = 0.0f;
                } else {
                    d = d * (y * (1.0f - y));
                }
                I_AT(d_inter, l, i, j) = d;
            }
        /* bias reversal: d_bs[l][j] += dZ[i][j] */
        for (int i = rows - 1; i >= 0; --i)
            for (int j = out_l - 1; j >= 0; --j)
                B_AT(d_bs, l, j) = B_AT(d_bs, l, j) + I_AT(d_inter, l, i, j);
        /* matmul reversal */
        int mrows = (l == 0) ? N : rows;
        for (int i = mrows - 1; i >= 0; --i)
            for (int j = out_l - 1; j >= 0; --j) {
                float dz = I_AT(d_inter, l, i, j);
                for (int k = in_l - 1; k >= 0; --k) {
                    float h = (l == 0) ? X[(size_t)i * C_in + k]
                                       : post[((size_t)(l - 1) * rows + i) * ld + k];
                    float w = W_AT(ws, l, k, j);
                    if (l == 0) {
                        if (d_X) d_X[(size_t)i * C_in + k] = d_X[(size_t)i * C_in + k] + dz * w;
                    } else {
                        I_AT(d_inter, l - 1, i, k) = I_AT(d_inter, l - 1, i, k) + dz * w;
                    }
                    W_AT(d_ws, l, k, j) = W_AT(d_ws, l, k, j) + h * dz;
                }
            }
    }
}

/* ---- full reverse of nerf_evaluate_and_march with seed g = _dreturn ------------------------
 * The reference's grad function re-runs the forward on its own tapes, so the primal scratch the
 * caller passes is left as it was; we therefore compute the forward privately.
 * Accumulating (+=) outputs: d_X [N][C_in], d_ws, d_bs, d_target [R][3], d_dists [R][S],
 * d_acc [R][3], d_inter [L][rows][ld].  (d_rgba, d_alpha, d_cumprod, d_weights end zero in the
 * reference and are not outputs here.) */
void oracle_nerf_backward(const float *X, int N, int C_in, const float *ws, const float *bs, int L,
                          const int *dims, int max_in, int max_out, const float *target, int R,
                          int S, const float *dists, int rows, int ld, float g, float *d_X,
                          float *d_ws, float *d_bs, float *d_target, float *d_dists, float *d_acc,
                          float *d_inter)
{
    size_t n_inter = (size_t)L * rows * ld, RS = (size_t)R * S;
    float *inter = (float *)calloc(n_inter, sizeof(float));
    float *rgba = (float *)calloc(RS * 4, sizeof(float));
    float *alpha = (float *)calloc(RS, sizeof(float));
    float *cum = (float *)calloc(RS, sizeof(float));
    float *wgt = (float *)calloc(RS, sizeof(float));
    float *acc = (float *)calloc((size_t)R * 3, sizeof(float));
    float *q = (float *)calloc(RS, sizeof(float));
    float *d_rgba = (float *)calloc(RS * 4, sizeof(float));
    float *d_alpha = (float *)calloc(RS, sizeof(float));
    float *d_cum = (float *)calloc(RS, sizeof(float));
    mlp_forward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_NERF);
    composite_forward(&I_AT(inter, L - 1, 0, 0), ld, R, S, dists, rgba, alpha, cum, wgt, acc);
    for (size_t i = 0; i < RS; ++i) q[i] = (1.0f - alpha[i]) + 1e-10f;

    /* loss reversal, nerf.py:297-302 */
    for (int r = R - 1; r >= 0; --r)
        for (int c = 2; c >= 0; --c) {
            float diff = acc[r * 3 + c] - target[r * 3 + c];
            float da = g * diff + diff * g;
            d_acc[r * 3 + c] = d_acc[r * 3 + c] + da;
            d_target[r * 3 + c] = d_target[r * 3 + c] - da;
        }
    /* accumulate reversal, nerf.py:281-288 : acc passes through, so d_acc keeps its value */
    for (int r = R - 1; r >= 0; --r) {
        float *dwv = (float *)calloc((size_t)S, sizeof(float)); /* adjoint of weights_samples */
        for (int s = S - 1; s >= 0; --s) {
            size_t i = (size_t)r * S + s;
            for (int c = 2; c >= 0; --c) {
                dwv[s] = dwv[s] + d_acc[r * 3 + c] * rgba[i * 4 + c];
                d_rgba[i * 4 + c] = d_rgba[i * 4 + c] + d_acc[r * 3 + c] * wgt[i];
            }
        }
        /* w = alpha * T reversal, nerf.py:267-272 (T_0 = 1, T_s = C_s) */
        for (int s = S - 1; s >= 0; --s) {
            size_t i = (size_t)r * S + s;
            d_alpha[i] = d_alpha[i] + dwv[s] * cum[i];
            d_cum[i] = d_cum[i] + dwv[s] * alpha[i];
        }
        free(dwv);
        /* cumprod[r][0] = 1 reversal, nerf.py:252-258 */
        d_cum[(size_t)r * S] = 0.0f;
        /* inclusive product reversal, nerf.py:226-232: C_s = C_{s-1} * q_s, s = S-1..1.
         * primal C_{s-1} at that point: the true inclusive product (C_0 = q_0 before the :=1). */
        {
            float *Ct = (float *)malloc((size_t)S * sizeof(float));
            Ct[0] = q[(size_t)r * S];
            for (int s = 1; s < S; ++s) Ct[s] = Ct[s - 1] * q[(size_t)r * S + s];
            for (int s = S - 1; s >= 1; --s) {
                size_t i = (size_t)r * S + s;
                float dC = d_cum[i];
                d_cum[i - 1] = d_cum[i - 1] + dC * q[i];
                d_cum[i] = dC * Ct[s - 1]; /* now the adjoint of q_s */
            }
            free(Ct);
        }
        /* q = 1 - alpha + 1e-10 reversal, nerf.py:215-220 */
        for (int s = S - 1; s >= 0; --s) {
            size_t i = (size_t)r * S + s;
            d_alpha[i] = d_alpha[i] - d_cum[i];
        }
        /* alpha = 1 - exp((0-sigma)*dist) reversal, nerf.py:200-205 */
        for (int s = S - 1; s >= 0; --s) {
            size_t i = (size_t)r * S + s;
            float sigma = rgba[i * 4 + 3];
            float e = expf((0.0f - sigma) * dists[i]);
            float darg = (0.0f - d_alpha[i]) * e;
            d_rgba[i * 4 + 3] = d_rgba[i * 4 + 3] - darg * dists[i];
            d_dists[i] = d_dists[i] + darg * (0.0f - sigma);
        }
    }
    /* copy reversal, nerf.py:182-191 */
    for (int r = R - 1; r >= 0; --r)
        for (int s = S - 1; s >= 0; --s)
            for (int c = 3; c >= 0; --c)
                I_AT(d_inter, L - 1, (size_t)r * S + s, c) =
                    I_AT(d_inter, L - 1, (size_t)r * S + s, c) + d_rgba[((size_t)r * S + s) * 4 + c];
    mlp_backward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_NERF, d_inter,
                 d_ws, d_bs, d_X);
    free(inter); free(rgba); free(alpha); free(cum); free(wgt); free(acc); free(q);
    free(d_rgba); free(d_alpha); free(d_cum);
}

/* reverse of mlp_fit (scripts/mlp_fit.py:1-147) with seed g */
void oracle_mlp_fit_backward(const float *X, int N, int C_in, const float *ws, const float *bs,
                             int L, const int *dims, int max_in, int max_out, const float *target,
                             int R, int Wt, int rows, int ld, float g, float *d_X, float *d_ws,
                             float *d_bs, float *d_target, float *d_inter)
{
    size_t n_inter = (size_t)L * rows * ld;
    float *inter = (float *)calloc(n_inter, sizeof(float));
    mlp_forward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_SIGMOID);
    for (int r = R - 1; r >= 0; --r)
        for (int c = Wt - 1; c >= 0; --c) {
            float diff = I_AT(inter, L - 1, r, c) - target[(size_t)r * Wt + c];
            float da = g * diff + diff * g;
            I_AT(d_inter, L - 1, r, c) = I_AT(d_inter, L - 1, r, c) + da;
            d_target[(size_t)r * Wt + c] = d_target[(size_t)r * Wt + c] - da;
        }
    mlp_backward(X, N, C_in, ws, bs, L, dims, max_in, max_out, rows, ld, inter, HEAD_SIGMOID,
                 d_inter, d_ws, d_bs, d_X);
    free(inter);
}

/* scripts/mlp_fit.py:150-172 ; c is accumulated into (Out array, caller zeroes) */
void oracle_mult_a_b(const float *a, int a_h, int a_w, const float *b, int b_h, int b_w, float *c)
{
    (void)b_h;
    for (int i = 0; i < a_h; ++i)
        for (int j = 0; j < b_w; ++j)
            for (int k = 0; k < a_w; ++k)
                c[(size_t)i * b_w + j] = c[(size_t)i * b_w + j] + a[(size_t)i * a_w + k] * b[(size_t)k * b_w + j];
}

/* pos_encoding.py:4-36 / 38-70: out[p][slot*F + f], slot 0 identity, slot 2i+1 = sin(2^i x),
 * slot 2i+2 = cos(2^i x); computed in float64, cast to float32 (pos_encoding.py:32,66). */
void oracle_pos_encoding(const double *x, long n_points, int F, int E, float *out)
{
    int C = F * (1 + 2 * E);
    for (long p = 0; p < n_points; ++p) {
        for (int f = 0; f < F; ++f) out[p * C + f] = (float)x[p * F + f];
        for (int i = 0; i < E; ++i) {
            double freq = ldexp(1.0, i);
            for (int f = 0; f < F; ++f) {
                out[p * C + (2 * i + 1) * F + f] = (float)sin(freq * x[p * F + f]);
                out[p * C + (2 * i + 2) * F + f] = (float)cos(freq * x[p * F + f]);
            }
        }
    }
}

/* train_nerf.py:289-311: pts[r][s] = o[r] + d[r] * t[r][s] in float64 (t given per ray/sample so
 * that both the reference's shared linspace and stratified samples are covered);
 * dists = [t[s+1]-t[s] ..., 1e8] cast to float32. */
void oracle_sample_points(const double *o, const double *d, const double *t, int R, int S,
                          double *pts, float *dists)
{
    for (int r = 0; r < R; ++r)
        for (int s = 0; s < S; ++s) {
            double tt = t[(size_t)r * S + s];
            for (int c = 0; c < 3; ++c)
                pts[((size_t)r * S + s) * 3 + c] = o[r * 3 + c] + d[r * 3 + c] * tt;
            dists[(size_t)r * S + s] =
                (s + 1 < S) ? (float)(t[(size_t)r * S + s + 1] - tt) : (float)1e8;
        }
}
