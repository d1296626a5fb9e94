/*
 * snn_oracle.c -- CPU restatement of the reference's spiking hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package
 * (snnimageclassification_b200/) may import, link or execute this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, and
 * only as the checker.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function
 * here against (i) the reference's own encoder known-answer tests and golden
 * image (test/test_to_spikes.py, test/test_x_to_spikes.npy, regenerated into
 * tests/golden/ by tests/golden/make_golden.py) and (ii) outputs of the
 * reference's PyTorch code itself (forward traces, loss, autograd gradients)
 * generated in the build container by the same script.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Arithmetic is fp32 with no FMA contraction except
 * where fmaf() is written explicitly; build with -ffp-contract=off.
 *
 * Summation orders (the reference leaves them to the BLAS; we fix them so the
 * CUDA kernels can be compared bit for bit):
 *   - input projection  : ascending k, one accumulator, fmaf
 *   - recurrent matvec  : sixteen accumulators over k mod 16, fmaf, balanced tree over adjacent accumulators
 *   - readout matvec    : ascending j, one accumulator
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int32_t B, T, N, H, O;
    int32_t layer_type; /* 0 = LIF, 1 = ALIF, 2 = Izhikevich  (spiking_layers.py:11-14) */
    int32_t surrogate;  /* 0 = FastSigmoid, 1 = Phi (spike_funcs.py:7-9) */
    int32_t recurrent;  /* use_recurrent_connection */
    float alpha, rho, theta, gamma, kappa, beta;
    /* Izhikevich (spiking_layers.py:275-296): dt, C, v_rest, v_th, k, a, b, c, d, v_peak */
    float dt, iz_C, iz_vr, iz_vth, iz_k, iz_a, iz_b, iz_c, iz_d, iz_vpeak;
} OracleCfg;

/* ------------------------------------------------------------------------ */
/* Encoder: datasets.py:42-54 (pixels_to_firing_periods)                     */
/* ------------------------------------------------------------------------ */

/* float64 pipeline (the reference's golden test runs in float64). */
static int64_t period_f64(double x, double t_max, double tau, double thr, double eps)
{
    int below = x < thr;                         /* datasets.py:49 */
    double lo = thr + eps;
    double xc = x < lo ? lo : (x > 1.0e9 ? 1.0e9 : x); /* :50 np.clip */
    double T = tau * log(xc / (xc - thr));       /* :51 */
    if (below) T = t_max;                        /* :52 */
    return (int64_t)T;                           /* :54 astype(int): trunc */
}

/* float32 pipeline: numpy keeps float32 when the python-float parameters meet a
 * float32 array (NEP 50 weak scalars), so thr, thr+eps, 1e9, tau and t_max are
 * all rounded to float32 first.  log is taken as the correctly rounded fp32
 * logarithm (fp64 log rounded once); numpy's SIMD logf can differ from it by
 * one ulp, which only matters within ~1e-5 of an integer latency -- never for
 * k/255 pixel levels (margin 5e-3, checked in tests). */
static int64_t period_f32(float x, double t_max, double tau, double thr, double eps)
{
    float thr_f = (float)thr;
    float lo = (float)(thr + eps);
    float hi = (float)1.0e9;
    int below = x < thr_f;
    float xc = x < lo ? lo : (x > hi ? hi : x);
    float d = xc - thr_f;
    float q = xc / d;
    float l = (float)log((double)q);
    float T = (float)tau * l;
    if (below) T = (float)t_max;
    return (int64_t)T;
}

int snn_oracle_periods_f64(const double* x, int64_t n, double t_max, double tau, double thr,
                           double eps, int64_t* out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = period_f64(x[i], t_max, tau, thr, eps);
    return 0;
}

int snn_oracle_periods_f32(const float* x, int64_t n, double t_max, double tau, double thr,
                           double eps, int64_t* out)
{
    for (int64_t i = 0; i < n; ++i) out[i] = period_f32(x[i], t_max, tau, thr, eps);
    return 0;
}

/* datasets.py:81-86 (firing_times_to_spikes) and :72-79
 * (firing_periods_to_spikes).  periods: (n_items, n_pix); out: (n_items,
 * n_steps, n_pix) uint8 in {0,1}. */
int snn_oracle_raster(const int64_t* periods, int64_t n_items, int64_t n_pix, int32_t n_steps,
                      int32_t periodic, uint8_t* out)
{
    memset(out, 0, (size_t)(n_items * n_steps * n_pix));
    for (int64_t it = 0; it < n_items; ++it) {
        for (int64_t p = 0; p < n_pix; ++p) {
            int64_t T = periods[it * n_pix + p];
            uint8_t* col = out + it * n_steps * n_pix + p;
            if (!periodic) {
                /* a negative latency would index from the end in numpy; the
                 * encoder never produces one (log(q) > 0 for q > 1). */
                if (T >= 0 && T < n_steps) col[T * n_pix] = 1;   /* :83-85 */
            } else {
                int64_t per = T;
                if (per > n_steps - 1) per = n_steps - 1;        /* :75 */
                if (per < 1) per = 1;                            /* :76 */
                for (int64_t t = per; t < n_steps; t += per)     /* :77-78 */
                    col[t * n_pix] = 1;
            }
        }
    }
    return 0;
}

int snn_oracle_encode_f64(const double* x, int64_t n_items, int64_t n_pix, int32_t n_steps,
                          double t_max, double tau, double thr, double eps, int32_t periodic,
                          uint8_t* out)
{
    int64_t n = n_items * n_pix;
    int64_t* per = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (!per) return -1;
    snn_oracle_periods_f64(x, n, t_max, tau, thr, eps, per);
    snn_oracle_raster(per, n_items, n_pix, n_steps, periodic, out);
    free(per);
    return 0;
}

int snn_oracle_encode_f32(const float* x, int64_t n_items, int64_t n_pix, int32_t n_steps,
                          double t_max, double tau, double thr, double eps, int32_t periodic,
                          uint8_t* out)
{
    int64_t n = n_items * n_pix;
    int64_t* per = (int64_t*)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (!per) return -1;
    snn_oracle_periods_f32(x, n, t_max, tau, thr, eps, per);
    snn_oracle_raster(per, n_items, n_pix, n_steps, periodic, out);
    free(per);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Forward: snn.py:201-219 unrolling spiking_layers.py:156-171 (LIF),        */
/* :229-243 (ALIF) and :402-408 (readout).                                   */
/* ------------------------------------------------------------------------ */

static float dot_rec16(const float* w_col, int stride, const float* z, int H)
{
    /* sum_k w[k*stride] * z[k]: sixteen accumulators over k mod 16 (the lanes of eight packed FFMA2 registers),
     * combined as a balanced tree over adjacent accumulators */
    float s[16];
    for (int j = 0; j < 16; ++j) s[j] = 0.f;
    for (int k = 0; k < H; k += 16)
        for (int j = 0; j < 16; ++j) s[j] = fmaf(w_col[(k + j) * stride], z[k + j], s[j]);
    float lo = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    float hi = ((s[8] + s[9]) + (s[10] + s[11])) + ((s[12] + s[13]) + (s[14] + s[15]));
    return lo + hi;
}

/* x (B,T,N); W_in (N,H); W_rec (H,H) raw, rec_mask (H,H) or NULL (= ones);
 * W_out (H,O); b_out (O).  V0/a0/Z0 (B,H) optional initial state (NULL =
 * zeros, spiking_layers.py:69-83).  Outputs (any may be NULL except V,Z,y):
 * I_in (B,T,H), V, a, Z (B,T,H), y (B,T,O). */
int snn_oracle_forward(const OracleCfg* c, const float* x, const float* W_in, const float* W_rec,
                       const float* rec_mask, const float* W_out, const float* b_out,
                       const float* V0, const float* a0, const float* Z0, float* I_in, float* V,
                       float* a, float* Z, float* y)
{
    const int B = c->B, T = c->T, N = c->N, H = c->H, O = c->O;
    if (H % 16) return -2;
    float* Weff = NULL;
    if (c->recurrent) {
        Weff = (float*)malloc(sizeof(float) * (size_t)H * H);
        if (!Weff) return -1;
        for (int i = 0; i < H * H; ++i)                      /* spiking_layers.py:165/235 */
            Weff[i] = rec_mask ? W_rec[i] * rec_mask[i] : W_rec[i];
    }
    float* vprev = (float*)calloc((size_t)H, sizeof(float));
    float* aprev = (float*)calloc((size_t)H, sizeof(float));
    float* zprev = (float*)calloc((size_t)H, sizeof(float));
    float* yprev = (float*)calloc((size_t)O, sizeof(float));
    float* cur = (float*)calloc((size_t)H, sizeof(float));
    if (!vprev || !aprev || !zprev || !yprev || !cur) return -1;

    for (int b = 0; b < B; ++b) {
        for (int i = 0; i < H; ++i) {
            vprev[i] = V0 ? V0[(size_t)b * H + i] : 0.f;
            aprev[i] = a0 ? a0[(size_t)b * H + i] : 0.f;
            zprev[i] = Z0 ? Z0[(size_t)b * H + i] : 0.f;
        }
        for (int o = 0; o < O; ++o) yprev[o] = 0.f;
        for (int t = 0; t < T; ++t) {
            const float* xt = x + ((size_t)b * T + t) * N;
            size_t row = ((size_t)b * T + t) * H;
            for (int i = 0; i < H; ++i) {
                float s = 0.f;                                   /* :163/233 x_t @ W_in */
                for (int k = 0; k < N; ++k) s = fmaf(xt[k], W_in[(size_t)k * H + i], s);
                cur[i] = s;
                if (I_in) I_in[row + i] = s;
            }
            if (c->layer_type == 2) {
                /* IzhikevichLayer.forward, spiking_layers.py:330-353; the trace `a` holds the recovery variable u */
                for (int i = 0; i < H; ++i) {
                    float I = c->recurrent ? cur[i] + dot_rec16(Weff + i, H, zprev, H) : cur[i] + 0.0f;   /* :344 */
                    float d1 = vprev[i] - c->iz_vr;
                    float d2 = vprev[i] - c->iz_vth;
                    float q = (c->iz_k * d1) * d2;               /* :345 */
                    q = q - aprev[i];
                    float dV = q + I;
                    float inc = (c->dt * dV) / c->iz_C;          /* :346 */
                    float v = (vprev[i] + inc) * (1.0f - zprev[i]);
                    v = v + c->iz_c * zprev[i];
                    float du = c->iz_a * (c->iz_b * d1 - aprev[i]);                                        /* :347 */
                    float un = (aprev[i] + c->dt * du) + c->iz_d * zprev[i];                               /* :348 */
                    float z = v >= c->iz_vpeak ? 1.0f : 0.0f;    /* :349 */
                    V[row + i] = v;
                    if (a) a[row + i] = un;
                    Z[row + i] = z;
                    vprev[i] = v;
                    aprev[i] = un;
                    cur[i] = z;
                }
            } else
            for (int i = 0; i < H; ++i) {
                float t1 = c->alpha * vprev[i];                  /* :169/239 */
                float t2 = t1 + cur[i];
                float t3 = c->recurrent ? t2 + dot_rec16(Weff + i, H, zprev, H) : t2 + 0.0f;
                float v = t3 * (1.0f - zprev[i]);
                float thr = c->theta;
                if (c->layer_type == 1) {
                    float an = c->rho * aprev[i];                /* :240 */
                    an = an + zprev[i];
                    float ba = c->beta * an;                     /* :241 */
                    thr = c->theta + ba;
                    if (a) a[row + i] = an;
                    aprev[i] = an;
                }
                float z = v >= thr ? 1.0f : 0.0f;                /* spike_funcs.py:27-28 */
                V[row + i] = v;
                Z[row + i] = z;
                vprev[i] = v;
                cur[i] = z; /* reuse: new spikes */
            }
            for (int i = 0; i < H; ++i) zprev[i] = cur[i];
            for (int o = 0; o < O; ++o) {                        /* spiking_layers.py:407 */
                float s = 0.f;
                for (int j = 0; j < H; ++j) s = fmaf(zprev[j], W_out[(size_t)j * O + o], s);
                float t1 = c->kappa * yprev[o];
                float t2 = t1 + s;
                float yo = t2 + b_out[o];
                y[((size_t)b * T + t) * O + o] = yo;
                yprev[o] = yo;
            }
        }
    }
    free(Weff); free(vprev); free(aprev); free(zprev); free(yprev); free(cur);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Head: snn.py:228 (max over time), :258 (log_softmax), :297 (NLLLoss mean) */
/* and its gradient w.r.t. the output trace.                                 */
/* ------------------------------------------------------------------------ */
int snn_oracle_head(int B, int T, int O, const float* y, const int64_t* labels, float* logits,
                    int32_t* tstar, float* logp, float* loss, float* g_y)
{
    double acc = 0.0;
    if (g_y) memset(g_y, 0, sizeof(float) * (size_t)B * T * O);
    for (int b = 0; b < B; ++b) {
        float lg[64];
        int ts[64];
        if (O > 64) return -2;
        for (int o = 0; o < O; ++o) {
            float m = y[((size_t)b * T) * O + o];
            int mt = 0;
            for (int t = 1; t < T; ++t) {
                float v = y[((size_t)b * T + t) * O + o];
                if (v > m) { m = v; mt = t; }          /* first max wins on ties */
            }
            lg[o] = m; ts[o] = mt;
            if (logits) logits[(size_t)b * O + o] = m;
            if (tstar) tstar[(size_t)b * O + o] = mt;
        }
        float mx = lg[0];
        for (int o = 1; o < O; ++o) if (lg[o] > mx) mx = lg[o];
        float se = 0.f;
        for (int o = 0; o < O; ++o) se += expf(lg[o] - mx);
        float lse = logf(se);
        for (int o = 0; o < O; ++o) {
            float lp = (lg[o] - mx) - lse;
            if (logp) logp[(size_t)b * O + o] = lp;
            if (labels) {
                if (o == (int)labels[b]) acc += -(double)lp;
                if (g_y) {
                    float g = (expf(lp) - (o == (int)labels[b] ? 1.0f : 0.0f)) / (float)B;
                    g_y[((size_t)b * T + ts[o]) * O + o] = g;
                }
            }
        }
    }
    if (loss) *loss = (float)(acc / (double)B);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* BPTT: the reverse sweep autograd performs over snn.py:209-214 for         */
/* batch_loss.backward() (snn.py:413), with the surrogate derivatives of     */
/* spike_funcs.py:59-62 (FastSigmoid) and :75-79 (Phi).  The threshold input */
/* receives no gradient (spike_funcs.py:62/79) and the reset is detached     */
/* (spiking_layers.py:169/239).                                              */
/* g_y (B,T,O): gradient w.r.t. the output trace.  g_Vs/g_Zs (B,T,H):        */
/* optional extra seeds on the hidden traces (NULL = none).                  */
/* Outputs: gI (B,T,H) optional; dW_in (N,H); dW_rec (H,H) masked; dW_out    */
/* (H,O); db (O).  Weight gradients are accumulated in double.               */
/* ------------------------------------------------------------------------ */
static float surrogate_grad(const OracleCfg* c, float v, float thr)
{
    if (c->surrogate == 0) {
        float d = c->gamma * fabsf(v - thr) + 1.0f;   /* spike_funcs.py:61 */
        return 1.0f / (d * d);
    } else {
        float te = thr + 1e-5f;                       /* spike_funcs.py:66,76-78 */
        float r = 1.0f - fabsf((v - thr) / te);
        if (r < 0.f) r = 0.f;
        return (c->gamma / te) * r;
    }
}

int snn_oracle_backward(const OracleCfg* c, const float* x, const float* W_rec,
                        const float* rec_mask, const float* W_out, const float* Z0,
                        const float* V, const float* a, const float* Z, const float* g_y,
                        const float* g_Vs, const float* g_Zs, float* gI, float* dW_in,
                        float* dW_rec, float* dW_out, float* db)
{
    const int B = c->B, T = c->T, N = c->N, H = c->H, O = c->O;
    double* aWin = (double*)calloc((size_t)N * H, sizeof(double));
    double* aWrec = (double*)calloc((size_t)H * H, sizeof(double));
    double* aWout = (double*)calloc((size_t)H * O, sizeof(double));
    double* adb = (double*)calloc((size_t)O, sizeof(double));
    float* Weff = (float*)calloc((size_t)H * H, sizeof(float));
    float* gy = (float*)calloc((size_t)O, sizeof(float));
    float* gv = (float*)calloc((size_t)H, sizeof(float));
    float* gu = (float*)calloc((size_t)H, sizeof(float));     /* Izhikevich: adjoint of the recovery variable */
    float* gi_next = (float*)calloc((size_t)H, sizeof(float));
    float* gi = (float*)calloc((size_t)H, sizeof(float));
    if (!aWin || !aWrec || !aWout || !adb || !Weff || !gy || !gv || !gi_next || !gi) return -1;
    if (c->recurrent)
        for (int i = 0; i < H * H; ++i) Weff[i] = rec_mask ? W_rec[i] * rec_mask[i] : W_rec[i];

    for (int b = 0; b < B; ++b) {
        for (int o = 0; o < O; ++o) gy[o] = 0.f;
        for (int i = 0; i < H; ++i) { gv[i] = 0.f; gi_next[i] = 0.f; if (gu) gu[i] = 0.f; }
        for (int t = T - 1; t >= 0; --t) {
            size_t row = ((size_t)b * T + t) * H;
            const float* zt = Z + row;
            const float* zp = t > 0 ? Z + row - H : (Z0 ? Z0 + (size_t)b * H : NULL);
            for (int o = 0; o < O; ++o) {
                gy[o] = g_y[((size_t)b * T + t) * O + o] + c->kappa * gy[o];
                adb[o] += gy[o];
            }
            for (int i = 0; i < H; ++i) {
                float s = 0.f;
                for (int o = 0; o < O; ++o) s = fmaf(gy[o], W_out[(size_t)i * O + o], s);
                if (c->recurrent) s += dot_rec16(Weff + (size_t)i * H, 1, gi_next, H);
                if (g_Zs) s += g_Zs[row + i];
                if (c->layer_type == 2) {
                    /* adjoint of spiking_layers.py:345-348 (reset factors detached, :343):
                     *   dV'/dV = (1 + dt k ((V-vr) + (V-vth)) / C)(1-Z)   dV'/du = -(dt/C)(1-Z)   dV'/dI = (dt/C)(1-Z)
                     *   du'/dV = dt a b                                    du'/du = 1 - dt a                          */
                    float v = V[row + i];
                    float sg = surrogate_grad(c, v, c->iz_vpeak);
                    float dq = c->iz_k * ((v - c->iz_vr) + (v - c->iz_vth));
                    float A = (1.0f + (c->dt * dq) / c->iz_C) * (1.0f - zt[i]);
                    float dtC = c->dt / c->iz_C;
                    float g = s * sg + gv[i] * A + gu[i] * (c->dt * c->iz_a * c->iz_b);
                    if (g_Vs) g += g_Vs[row + i];
                    float gun = gv[i] * (-dtC) * (1.0f - zt[i]) + gu[i] * (1.0f - c->dt * c->iz_a);
                    gv[i] = g;
                    gu[i] = gun;
                    gi[i] = (g * dtC) * (1.0f - (zp ? zp[i] : 0.f));
                } else {
                float thr = c->theta;
                if (c->layer_type == 1) thr = c->theta + c->beta * a[row + i];
                float sg = surrogate_grad(c, V[row + i], thr);
                float carry = c->alpha * gv[i] * (1.0f - zt[i]);
                float g = s * sg + carry;
                if (g_Vs) g += g_Vs[row + i];
                gv[i] = g;
                gi[i] = g * (1.0f - (zp ? zp[i] : 0.f));
                }
                if (gI) gI[row + i] = gi[i];
                for (int o = 0; o < O; ++o) aWout[(size_t)i * O + o] += (double)zt[i] * gy[o];
            }
            const float* xt = x + ((size_t)b * T + t) * N;
            for (int k = 0; k < N; ++k) {
                if (xt[k] == 0.f) continue;
                for (int i = 0; i < H; ++i) aWin[(size_t)k * H + i] += (double)xt[k] * gi[i];
            }
            if (c->recurrent && zp) {
                for (int j = 0; j < H; ++j) {
                    if (zp[j] == 0.f) continue;
                    for (int i = 0; i < H; ++i) aWrec[(size_t)j * H + i] += (double)zp[j] * gi[i];
                }
            }
            for (int i = 0; i < H; ++i) gi_next[i] = gi[i];
        }
    }
    for (int i = 0; i < N * H; ++i) dW_in[i] = (float)aWin[i];
    if (dW_rec)
        for (int i = 0; i < H * H; ++i)
            dW_rec[i] = (float)(rec_mask ? aWrec[i] * rec_mask[i] : aWrec[i]);
    for (int i = 0; i < H * O; ++i) dW_out[i] = (float)aWout[i];
    for (int o = 0; o < O; ++o) db[o] = (float)adb[o];
    free(aWin); free(aWrec); free(aWout); free(adb); free(Weff); free(gy); free(gv);
    free(gi_next); free(gi); free(gu);
    return 0;
}
