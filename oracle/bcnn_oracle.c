/*
 * bcnn_oracle.c -- TEST INFRASTRUCTURE ONLY (see bcnn_oracle.h).
 *
 * CPU restatement of bcnn's CPU layer arithmetic, written from the behaviour of
 * the reference (file:line cited per function, relative to /root/reference).
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction: the reference's AVX
 * kernels use separate mul and add, src/kernels/bcnn_mat.c:2307-2352).
 */
#include "bcnn_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ */
/* Reductions with the reference's SSE lane order                      */
/* ------------------------------------------------------------------ */

/* bcnn_vsum, src/kernels/bcnn_mat.c:447-475 (AVX build): one 4-lane accumulator,
 * two adds per 8 elements, horizontal ((l0+l1)+l2)+l3, then a scalar tail. */
float orc_vsum(int n, const float *x) {
    float lane[4] = {0.f, 0.f, 0.f, 0.f};
    int nd = n / 8 * 8;
    for (int i = 0; i < nd; i += 8) {
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + x[i + j];
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + x[i + 4 + j];
    }
    float s = 0.f;
    s += lane[0] + lane[1] + lane[2] + lane[3];
    for (int i = nd; i < n; ++i) s += x[i];
    return s;
}

/* bcnn_dot, src/kernels/bcnn_mat.c:413-445. */
float orc_dot(int n, const float *x, const float *y) {
    float lane[4] = {0.f, 0.f, 0.f, 0.f};
    int nd = n / 8 * 8;
    for (int i = 0; i < nd; i += 8) {
        float p0[4], p1[4];
        for (int j = 0; j < 4; ++j) p0[j] = x[i + j] * y[i + j];
        for (int j = 0; j < 4; ++j) p1[j] = x[i + 4 + j] * y[i + 4 + j];
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + p0[j];
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + p1[j];
    }
    float s = 0.f;
    s += lane[0] + lane[1] + lane[2] + lane[3];
    for (int i = nd; i < n; ++i) s += x[i] * y[i];
    return s;
}

/* bcnn_shiftdot, src/kernels/bcnn_mat.c:652-690. */
float orc_shiftdot(int n, const float *x, float a, const float *y, float b) {
    float lane[4] = {0.f, 0.f, 0.f, 0.f};
    int nd = n / 8 * 8;
    for (int i = 0; i < nd; i += 8) {
        float p0[4], p1[4];
        for (int j = 0; j < 4; ++j) p0[j] = (x[i + j] - a) * (y[i + j] - b);
        for (int j = 0; j < 4; ++j)
            p1[j] = (x[i + 4 + j] - a) * (y[i + 4 + j] - b);
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + p0[j];
        for (int j = 0; j < 4; ++j) lane[j] = lane[j] + p1[j];
    }
    float s = 0.f;
    s += lane[0] + lane[1] + lane[2] + lane[3];
    for (int i = nd; i < n; ++i) s += (x[i] - a) * (y[i] - b);
    return s;
}

/* ------------------------------------------------------------------ */
/* Convolution                                                          */
/* ------------------------------------------------------------------ */

/* Output extent, src/layers/bcnn_conv_layer.c:126-134. */
int orc_conv_out_dim(int in, int k, int stride, int pad) {
    return (in + 2 * pad - k) / stride + 1;
}

/* Value the reference's im2col would place at row (c,kh,kw), column (ho,wo):
 * src/kernels/bcnn_mat.c:817-854 (zero outside the image). */
static inline float im2col_at(const float *img, int h, int wd, int c, int kh,
                              int kw, int ho, int wo, int stride, int pad) {
    int ih = ho * stride - pad + kh;
    int iw = wo * stride - pad + kw;
    if ((unsigned)ih >= (unsigned)h || (unsigned)iw >= (unsigned)wd) return 0.f;
    return img[(c * h + ih) * wd + iw];
}

/* bcnn_forward_conv_layer_cpu, src/layers/bcnn_conv_layer.c:438-462: for every
 * image b and group g, C[m x n] += A[m x k] * im2col(src)[k x n] with
 * m = cout/groups, k = (cin/groups)*ksz*ksz, n = ho*wo; dst zero-filled first
 * (:389). The K sum runs in (c,kh,kw) order as in the GEMM's K loop. */
void orc_conv_forward(const float *x, const float *w, float *y, int n, int cin,
                      int h, int wd, int cout, int k, int stride, int pad,
                      int groups) {
    int ho = orc_conv_out_dim(h, k, stride, pad);
    int wo = orc_conv_out_dim(wd, k, stride, pad);
    int cg = cin / groups, mg = cout / groups;
#pragma omp parallel for collapse(2)
    for (int b = 0; b < n; ++b) {
        for (int co = 0; co < cout; ++co) {
            int g = co / mg;
            const float *img = x + ((size_t)b * cin + (size_t)g * cg) * h * wd;
            const float *wrow = w + (size_t)co * cg * k * k;
            float *out = y + ((size_t)b * cout + co) * ho * wo;
            for (int oh = 0; oh < ho; ++oh)
                for (int ow = 0; ow < wo; ++ow) {
                    float acc = 0.f;
                    for (int c = 0; c < cg; ++c)
                        for (int kh = 0; kh < k; ++kh)
                            for (int kw = 0; kw < k; ++kw)
                                acc += wrow[(c * k + kh) * k + kw] *
                                       im2col_at(img, h, wd, c, kh, kw, oh, ow,
                                                 stride, pad);
                    out[oh * wo + ow] = acc;
                }
        }
    }
}

/* bcnn_backward_conv_layer_cpu, src/layers/bcnn_conv_layer.c:534-583.
 * Weight gradient: gW[m x kk] += dy[m x n] * col^T (beta = 1, :551), images
 * visited in batch order. Data gradient: col = W^T * dy (beta = 0, :575) then
 * bcnn_col2im, which zero-fills the image first (bcnn_mat.c:944) => dx is
 * overwritten, not accumulated. */
void orc_conv_backward(const float *x, const float *w, const float *dy,
                       float *gw, float *dx, int n, int cin, int h, int wd,
                       int cout, int k, int stride, int pad, int groups) {
    int ho = orc_conv_out_dim(h, k, stride, pad);
    int wo = orc_conv_out_dim(wd, k, stride, pad);
    int cg = cin / groups, mg = cout / groups;
    /* wgrad */
#pragma omp parallel for
    for (int co = 0; co < cout; ++co) {
        int g = co / mg;
        for (int c = 0; c < cg; ++c)
            for (int kh = 0; kh < k; ++kh)
                for (int kw = 0; kw < k; ++kw) {
                    float *dst = gw + (((size_t)co * cg + c) * k + kh) * k + kw;
                    for (int b = 0; b < n; ++b) {
                        const float *img =
                            x + ((size_t)b * cin + (size_t)g * cg) * h * wd;
                        const float *g_out =
                            dy + ((size_t)b * cout + co) * ho * wo;
                        float acc = 0.f;
                        for (int oh = 0; oh < ho; ++oh)
                            for (int ow = 0; ow < wo; ++ow)
                                acc += g_out[oh * wo + ow] *
                                       im2col_at(img, h, wd, c, kh, kw, oh, ow,
                                                 stride, pad);
                        *dst += acc;
                    }
                }
    }
    if (!dx) return;
    /* dgrad (gather form of GEMM + col2im) */
#pragma omp parallel for collapse(2)
    for (int b = 0; b < n; ++b) {
        for (int ci = 0; ci < cin; ++ci) {
            int g = ci / cg, c = ci % cg;
            float *out = dx + ((size_t)b * cin + ci) * h * wd;
            for (int ih = 0; ih < h; ++ih)
                for (int iw = 0; iw < wd; ++iw) {
                    float acc = 0.f;
                    for (int kh = 0; kh < k; ++kh) {
                        int th = ih + pad - kh;
                        if (th < 0 || th % stride) continue;
                        int oh = th / stride;
                        if (oh >= ho) continue;
                        for (int kw = 0; kw < k; ++kw) {
                            int tw = iw + pad - kw;
                            if (tw < 0 || tw % stride) continue;
                            int ow = tw / stride;
                            if (ow >= wo) continue;
                            float col = 0.f; /* col[(c,kh,kw),(oh,ow)] = sum_m W^T dy */
                            for (int m = 0; m < mg; ++m) {
                                int co = g * mg + m;
                                col += w[(((size_t)co * cg + c) * k + kh) * k +
                                         kw] *
                                       dy[(((size_t)b * cout + co) * ho + oh) *
                                              wo +
                                          ow];
                            }
                            acc += col;
                        }
                    }
                    out[ih * wd + iw] = acc;
                }
        }
    }
}

/* bcnn_add_bias, src/kernels/bcnn_mat.c:761-770. (The AVX bcnn_add_scalar skips
 * a bias of exactly 1.0f, :368-411 -- synthetic biases avoid that value.) */
void orc_add_bias(float *y, const float *b, int n, int c, int hw) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < c; ++j) {
            float *p = y + ((size_t)i * c + j) * hw;
            for (int s = 0; s < hw; ++s) p[s] += b[j];
        }
}

/* bcnn_grad_bias, src/kernels/bcnn_mat.c:798-811: sequential += into gb[c],
 * batch-major. */
void orc_grad_bias(float *gb, const float *dy, int n, int c, int hw) {
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < c; ++j) {
            const float *p = dy + ((size_t)i * c + j) * hw;
            for (int s = 0; s < hw; ++s) gb[j] += p[s];
        }
}

/* ------------------------------------------------------------------ */
/* Batchnorm                                                            */
/* ------------------------------------------------------------------ */

/* bcnn_forward_batchnorm_cpu, src/layers/bcnn_batchnorm_layer.c:196-242 with
 * _mean_variance_forward (:147-168), _norm_forward (:170-182, eps 1e-6),
 * scale_and_add_bias (:184-194), bcnn_scales / bcnn_add_bias. */
void orc_bn_forward(float *y, int n, int c, int hw, float *run_mean,
                    float *run_var, const float *gamma, const float *beta,
                    float *saved_mean, float *saved_var, float *x_norm,
                    float *x_copy, int mode) {
    size_t total = (size_t)n * c * hw;
    if (mode == ORC_MODE_PREDICT) { /* stats are pre-folded into gamma/beta */
        for (int b = 0; b < n; ++b)
            for (int j = 0; j < c; ++j) {
                float *p = y + ((size_t)b * c + j) * hw;
                for (int i = 0; i < hw; ++i) p[i] = p[i] * gamma[j] + beta[j];
            }
        return;
    }
    if (x_copy) memcpy(x_copy, y, total * sizeof(float));
    const float *mean = run_mean, *var = run_var;
    if (mode == ORC_MODE_TRAIN) {
        float scale = 1.0f / (n * hw);
        for (int j = 0; j < c; ++j) {
            float m = 0.f, v = 0.f;
            for (int b = 0; b < n; ++b) {
                const float *p = y + ((size_t)b * c + j) * hw;
                m += orc_vsum(hw, p);
                v += orc_dot(hw, p, p);
            }
            saved_mean[j] = m;
            saved_var[j] = v;
        }
        for (int j = 0; j < c; ++j) saved_mean[j] *= scale; /* bcnn_scal */
        for (int j = 0; j < c; ++j)                         /* bcnn_varmean */
            saved_var[j] = saved_var[j] * scale - saved_mean[j] * saved_mean[j];
        for (int j = 0; j < c; ++j) { /* running = .9 running + .1 batch (:221-224) */
            run_mean[j] *= 0.9f;
            run_mean[j] += 0.1f * saved_mean[j];
            run_var[j] *= 0.9f;
            run_var[j] += 0.1f * saved_var[j];
        }
        mean = saved_mean;
        var = saved_var;
    }
    for (int b = 0; b < n; ++b)
        for (int j = 0; j < c; ++j) {
            float *p = y + ((size_t)b * c + j) * hw;
            for (int i = 0; i < hw; ++i)
                p[i] = (p[i] - mean[j]) / (sqrtf(var[j] + 0.000001f));
        }
    if (mode == ORC_MODE_TRAIN && x_norm) memcpy(x_norm, y, total * sizeof(float));
    for (int b = 0; b < n; ++b) /* bcnn_scales then bcnn_add_bias */
        for (int j = 0; j < c; ++j) {
            float *p = y + ((size_t)b * c + j) * hw;
            if (gamma[j] == 0.0f) { /* bcnn_scal: a==0 -> memset, a==1 -> no-op */
                for (int i = 0; i < hw; ++i) p[i] = 0.f;
            } else if (gamma[j] != 1.0f) {
                for (int i = 0; i < hw; ++i) p[i] *= gamma[j];
            }
            for (int i = 0; i < hw; ++i) p[i] += beta[j];
        }
}

/* bcnn_backward_batchnorm_cpu, src/layers/bcnn_batchnorm_layer.c:301-332 with
 * _mean_variance_backward (:263-281), bcnn_varnorm (bcnn_mat.c:692-727) and
 * _normalize_backward (:283-299). eps is 1e-5 here (1e-6 in forward). */
void orc_bn_backward(float *dy, int n, int c, int hw, const float *gamma,
                     float *g_gamma, float *g_beta, const float *saved_mean,
                     const float *saved_var, float *d_mean, float *d_var,
                     const float *x_norm, const float *x_copy) {
    orc_grad_bias(g_beta, dy, n, c, hw);
    for (int j = 0; j < c; ++j) { /* bcnn_grad_scales, bcnn_mat.c:783-796 */
        float s = 0.f;
        for (int b = 0; b < n; ++b) {
            size_t off = ((size_t)b * c + j) * hw;
            for (int i = 0; i < hw; ++i) s += dy[off + i] * x_norm[off + i];
        }
        g_gamma[j] += s;
    }
    for (int b = 0; b < n; ++b) /* dy *= gamma (bcnn_scales) */
        for (int j = 0; j < c; ++j) {
            float *p = dy + ((size_t)b * c + j) * hw;
            if (gamma[j] == 0.0f) {
                for (int i = 0; i < hw; ++i) p[i] = 0.f;
            } else if (gamma[j] != 1.0f) {
                for (int i = 0; i < hw; ++i) p[i] *= gamma[j];
            }
        }
    for (int j = 0; j < c; ++j) {
        float dm = 0.f, dv = 0.f;
        for (int b = 0; b < n; ++b) {
            size_t off = ((size_t)b * c + j) * hw;
            dm += orc_vsum(hw, dy + off);
            dv += orc_shiftdot(hw, x_copy + off, saved_mean[j], dy + off, 0.0f);
        }
        d_mean[j] = dm * (-1.0f / sqrtf(saved_var[j] + 0.00001f));
        d_var[j] = dv;
    }
    for (int j = 0; j < c; ++j)
        d_var[j] *= -0.5f / (saved_var[j] * sqrtf(saved_var[j]) + 0.00001f);
    for (int b = 0; b < n; ++b)
        for (int j = 0; j < c; ++j) {
            size_t off = ((size_t)b * c + j) * hw;
            for (int i = 0; i < hw; ++i)
                dy[off + i] =
                    dy[off + i] * 1.0f / (sqrtf(saved_var[j] + 0.00001f)) +
                    d_var[j] * 2.0f * (x_copy[off + i] - saved_mean[j]) /
                        (hw * n) +
                    d_mean[j] / (hw * n);
        }
}

/* ------------------------------------------------------------------ */
/* Activations                                                          */
/* ------------------------------------------------------------------ */

/* bcnn_forward_activation_cpu, src/layers/bcnn_activation_layer.c:90-146. */
void orc_activation_forward(float *x, int sz, const float *slope, int hw, int c,
                            int act) {
    switch (act) {
        case ORC_ACT_TANH:
            for (int i = 0; i < sz; ++i)
                x[i] = (float)(exp(2 * x[i]) - 1) / ((float)exp(2 * x[i]) + 1);
            break;
        case ORC_ACT_RELU:
            for (int i = 0; i < sz; ++i) x[i] = x[i] * (x[i] > 0);
            break;
        case ORC_ACT_LRELU: /* slope is 0.1 (:106), not the header's 0.01 */
            for (int i = 0; i < sz; ++i) x[i] = (x[i] > 0 ? x[i] : 0.1f * x[i]);
            break;
        case ORC_ACT_RAMP:
            for (int i = 0; i < sz; ++i) x[i] = x[i] * (x[i] > 0) + 0.1f * x[i];
            break;
        case ORC_ACT_SOFTPLUS:
            for (int i = 0; i < sz; ++i)
                x[i] = (float)log(1.0f + (float)exp(x[i]));
            break;
        case ORC_ACT_ABS:
            for (int i = 0; i < sz; ++i) x[i] = (float)fabs(x[i]);
            break;
        case ORC_ACT_CLAMP:
            for (int i = 0; i < sz; ++i)
                x[i] = x[i] < 0.f ? 0.f : (x[i] > 1.f ? 1.f : x[i]);
            break;
        case ORC_ACT_LOGISTIC:
            for (int i = 0; i < sz; ++i)
                x[i] = 1.0f / (1.0f + (float)exp(-x[i]));
            break;
        case ORC_ACT_PRELU:
            for (int i = 0; i < sz; ++i) {
                int ch = (i / hw) % c;
                x[i] = (x[i] > 0 ? x[i] : slope[ch] * x[i]);
            }
            break;
        default:
            break;
    }
}

/* bcnn_backward_activation_cpu, src/layers/bcnn_activation_layer.c:165-226:
 * derivative evaluated on the post-activation value y. */
void orc_activation_backward(const float *y, float *dy, int sz,
                             const float *slope, float *g_slope, int hw, int c,
                             int act) {
    switch (act) {
        case ORC_ACT_TANH:
            for (int i = 0; i < sz; ++i) dy[i] *= (1 - y[i] * y[i]);
            break;
        case ORC_ACT_RELU:
            for (int i = 0; i < sz; ++i) dy[i] *= ((float)(y[i] > 0));
            break;
        case ORC_ACT_LRELU:
            for (int i = 0; i < sz; ++i) dy[i] *= (y[i] > 0 ? 1.0f : 0.1f);
            break;
        case ORC_ACT_RAMP:
            for (int i = 0; i < sz; ++i) dy[i] *= ((float)(y[i] > 0) + 0.1f);
            break;
        case ORC_ACT_SOFTPLUS:
            for (int i = 0; i < sz; ++i)
                dy[i] *= 1.0f / (1.0f + (float)exp(-y[i]));
            break;
        case ORC_ACT_ABS:
            for (int i = 0; i < sz; ++i) dy[i] *= (y[i] >= 0 ? 1.0f : -1.0f);
            break;
        case ORC_ACT_CLAMP:
            for (int i = 0; i < sz; ++i)
                dy[i] *= ((float)(y[i] > 0.0f && y[i] < 1.0f));
            break;
        case ORC_ACT_LOGISTIC:
            for (int i = 0; i < sz; ++i) dy[i] *= (1 - y[i]) * y[i];
            break;
        case ORC_ACT_PRELU:
            for (int i = 0; i < sz; ++i) {
                int ch = (i / hw) % c;
                g_slope[ch] += dy[i] * y[i] * (y[i] < 0);
            }
            for (int i = 0; i < sz; ++i) {
                int ch = (i / hw) % c;
                dy[i] *= (y[i] > 0 ? 1.0f : slope[ch]);
            }
            break;
        default:
            break;
    }
}

/* ------------------------------------------------------------------ */
/* Pooling                                                              */
/* ------------------------------------------------------------------ */

/* Output extent per padding policy, src/layers/bcnn_maxpool_layer.c:62-83. */
int orc_maxpool_out_dim(int in, int k, int stride, int padding) {
    if (padding == ORC_PAD_SAME) return (in + stride - 1) / stride;
    if (padding == ORC_PAD_VALID) return (in - k + stride) / stride;
    if (padding == ORC_PAD_CAFFE)
        return (int)(ceil((float)(in - k) / stride)) + 1;
    return 0;
}

/* bcnn_forward_maxpool_layer_cpu, src/layers/bcnn_maxpool_layer.c:145-191:
 * window starts at (i*stride, j*stride) with no leading pad; out-of-image taps
 * read -FLT_MAX; strict '>' so the first maximum in row-major window order
 * wins; the recorded index is the flat NCHW offset including the batch, or -1
 * when nothing exceeds -FLT_MAX. */
void orc_maxpool_forward(const float *x, float *y, int *idx, int n, int c, int h,
                         int w, int k, int stride, int ho, int wo) {
    for (int b = 0; b < n; ++b)
        for (int ch = 0; ch < c; ++ch)
            for (int i = 0; i < ho; ++i)
                for (int j = 0; j < wo; ++j) {
                    int dst = j + wo * (i + ho * (ch + c * b));
                    float best = -FLT_MAX;
                    int best_i = -1;
                    for (int r = 0; r < k; ++r)
                        for (int s = 0; s < k; ++s) {
                            int ih = i * stride + r, iw = j * stride + s;
                            int src = iw + w * (ih + h * (ch + b * c));
                            int ok = (ih >= 0 && ih < h && iw >= 0 && iw < w);
                            float v = ok ? x[src] : -FLT_MAX;
                            if (v > best) {
                                best = v;
                                best_i = src;
                            }
                        }
                    y[dst] = best;
                    idx[dst] = best_i;
                }
}

/* bcnn_backward_maxpool_layer_cpu, src/layers/bcnn_maxpool_layer.c:258-273:
 * scatter-add. (The reference would write grad[-1] for an index of -1; that is
 * undefined behaviour there, skipped here.) */
void orc_maxpool_backward(float *dx, const float *dy, const int *idx,
                          int out_sz) {
    for (int i = 0; i < out_sz; ++i)
        if (idx[i] >= 0) dx[idx[i]] += dy[i];
}

/* bcnn_forward_avgpool_layer_cpu, src/layers/bcnn_avgpool_layer.c:82-99:
 * sequential sum then one divide. */
void orc_avgpool_forward(const float *x, float *y, int n, int c, int hw) {
    for (int i = 0; i < n * c; ++i) {
        float s = 0;
        for (int j = 0; j < hw; ++j) s += x[(size_t)i * hw + j];
        y[i] = s / hw;
    }
}

/* bcnn_backward_avgpool_layer_cpu, src/layers/bcnn_avgpool_layer.c:109-125. */
void orc_avgpool_backward(float *dx, const float *dy, int n, int c, int hw) {
    for (int i = 0; i < n * c; ++i)
        for (int j = 0; j < hw; ++j) dx[(size_t)i * hw + j] += dy[i] / hw;
}

/* ------------------------------------------------------------------ */
/* Depthwise convolution                                                */
/* ------------------------------------------------------------------ */

/* bcnn_forward_depthwise_conv_layer_cpu, src/layers/bcnn_depthwise_conv_layer.c
 * :165-293: direct k x k correlation per channel, zero padding, taps summed in
 * (kh,kw) order. Bias and activation are separate (orc_add_bias,
 * orc_activation_forward). */
void orc_depthwise_forward(const float *x, const float *w, float *y, int n,
                           int c, int h, int wd, int k, int stride, int pad) {
    int ho = orc_conv_out_dim(h, k, stride, pad);
    int wo = orc_conv_out_dim(wd, k, stride, pad);
    for (int b = 0; b < n; ++b)
        for (int ch = 0; ch < c; ++ch) {
            const float *img = x + ((size_t)b * c + ch) * h * wd;
            const float *wk = w + (size_t)ch * k * k;
            float *out = y + ((size_t)b * c + ch) * ho * wo;
            for (int oh = 0; oh < ho; ++oh)
                for (int ow = 0; ow < wo; ++ow) {
                    float v = 0;
                    for (int kh = 0; kh < k; ++kh)
                        for (int kw = 0; kw < k; ++kw) {
                            int ih = -pad + oh * stride + kh;
                            int iw = -pad + ow * stride + kw;
                            if (ih >= 0 && ih < h && iw >= 0 && iw < wd)
                                v += wk[kh * k + kw] * img[ih * wd + iw];
                        }
                    out[oh * wo + ow] = v;
                }
        }
}

/* bcnn_backward_depthwise_conv_layer_cpu, src/layers/
 * bcnn_depthwise_conv_layer.c:295-547: both gw and dx accumulate with '+=', in
 * n -> c -> h -> w -> kh -> kw order; both are computed only when the source
 * has a gradient buffer (:318), which the caller expresses by passing dx. */
void orc_depthwise_backward(const float *x, const float *w, const float *dy,
                            float *gw, float *dx, int n, int c, int h, int wd,
                            int k, int stride, int pad) {
    if (!dx) return;
    int ho = orc_conv_out_dim(h, k, stride, pad);
    int wo = orc_conv_out_dim(wd, k, stride, pad);
    for (int b = 0; b < n; ++b)
        for (int ch = 0; ch < c; ++ch) {
            const float *img = x + ((size_t)b * c + ch) * h * wd;
            float *gimg = dx + ((size_t)b * c + ch) * h * wd;
            const float *g_out = dy + ((size_t)b * c + ch) * ho * wo;
            for (int oh = 0; oh < ho; ++oh)
                for (int ow = 0; ow < wo; ++ow)
                    for (int kh = 0; kh < k; ++kh)
                        for (int kw = 0; kw < k; ++kw) {
                            int ih = -pad + oh * stride + kh;
                            int iw = -pad + ow * stride + kw;
                            if (ih >= 0 && ih < h && iw >= 0 && iw < wd) {
                                gw[(size_t)ch * k * k + kh * k + kw] +=
                                    img[ih * wd + iw] * g_out[oh * wo + ow];
                                gimg[ih * wd + iw] +=
                                    w[(size_t)ch * k * k + kh * k + kw] *
                                    g_out[oh * wo + ow];
                            }
                        }
        }
}

/* ------------------------------------------------------------------ */
/* Optimizer                                                            */
/* ------------------------------------------------------------------ */

/* bcnn_sgd_update_cpu, src/bcnn_learner.c:67-83: momentum lives in the gradient
 * buffer (grad *= momentum after the step; the next backward accumulates on top). */
void orc_sgd_update(float *w, float *b, float *gw, float *gb, int wsz, int bsz,
                    int batch, float lr, float momentum, float decay) {
    if (b && gb) {
        float a = -lr / batch;
        for (int i = 0; i < bsz; ++i) b[i] += a * gb[i];
        for (int i = 0; i < bsz; ++i) gb[i] *= momentum;
    }
    if (w && gw) {
        float d = decay * batch, a = -lr / batch;
        for (int i = 0; i < wsz; ++i) gw[i] += d * w[i];
        for (int i = 0; i < wsz; ++i) w[i] += a * gw[i];
        for (int i = 0; i < wsz; ++i) gw[i] *= momentum;
    }
}

/* bcnn_adam_update_cpu, src/bcnn_learner.c:106-132, pass by pass. `iter` is what the layers
 * hand in: learner->seen, samples not steps (bcnn_conv_layer.c:831). The division is bcnn_vdiv
 * as the default (AVX) build compiles it, src/kernels/bcnn_mat.c:277-310: the 8-wide body
 * divides unconditionally, the n % 8 scalar tail yields 0 when |denominator| <= 1e-5. */
void orc_adam_update(float *w, float *b, float *gw, float *gb, float *m, float *v, int wsz,
                     int bsz, int batch, int iter, float beta1, float beta2, float lr,
                     float momentum, float decay) {
    float mu = sqrtf(1.0f - powf(beta2, (float)iter + 1)) / (1.0f - powf(beta1, (float)iter + 1));
    if (b && gb) {
        float a = -lr / batch;
        for (int i = 0; i < bsz; ++i) b[i] += a * gb[i];
        for (int i = 0; i < bsz; ++i) gb[i] *= momentum;
    }
    if (w && gw) {
        float d = decay * batch, a = -lr / batch * mu;
        int body = wsz / 8 * 8;
        for (int i = 0; i < wsz; ++i) gw[i] += d * w[i];
        for (int i = 0; i < wsz; ++i) m[i] = (1.0f - beta1) * gw[i] + beta1 * m[i];
        for (int i = 0; i < wsz; ++i) gw[i] = gw[i] * gw[i];
        for (int i = 0; i < wsz; ++i) v[i] = (1.0f - beta2) * gw[i] + beta2 * v[i];
        for (int i = 0; i < wsz; ++i) gw[i] = powf(v[i], 0.5f);
        for (int i = 0; i < wsz; ++i) gw[i] += 0.0000001f;
        for (int i = 0; i < body; ++i) gw[i] = m[i] / gw[i];
        for (int i = body; i < wsz; ++i) gw[i] = fabsf(gw[i]) > 0.00001f ? m[i] / gw[i] : 0.0f;
        for (int i = 0; i < wsz; ++i) w[i] += a * gw[i];
        memset(gw, 0, (size_t)wsz * sizeof(float));
    }
}

/* Inference part of bcnn_forward_yolo_layer_cpu, src/layers/bcnn_yolo.c:226-250: copy, then per
 * sample and anchor group the logistic function (bcnn_activation_layer.c: 1/(1+exp(-x)), exp in
 * double) on entries 0..1 and coords..coords+classes; entry_index of :207-215 spelled out. */
void orc_yolo_forward(const float *x, float *y, int n, int num, int classes, int coords, int hw) {
    int group = coords + classes + 1, chw = num * group * hw;
    memcpy(y, x, (size_t)n * chw * sizeof(float));
    for (int b = 0; b < n; ++b)
        for (int a = 0; a < num; ++a) {
            float *p = y + (size_t)b * chw + (size_t)a * hw * group;
            for (int i = 0; i < 2 * hw; ++i) p[i] = 1.0f / (1.0f + (float)exp(-p[i]));
            p += (size_t)coords * hw;
            for (int i = 0; i < (1 + classes) * hw; ++i) p[i] = 1.0f / (1.0f + (float)exp(-p[i]));
        }
}

/* ------------------------------------------------------------------ */
/* Glue: fc, softmax, eltwise                                           */
/* ------------------------------------------------------------------ */

/* bcnn_forward_fullc_layer_cpu, src/layers/bcnn_fc_layer.c:144-181: per output a
 * sum over input channels of bcnn_dot over the spatial extent, then + bias. */
void orc_fc_forward(const float *x, const float *w, const float *b, float *y,
                    int n, int in_c, int in_hw, int out) {
    int in_sz = in_c * in_hw;
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < out; ++p) {
            float s = 0.f;
            for (int q = 0; q < in_c; ++q)
                s += orc_dot(in_hw, x + (size_t)i * in_sz + q * in_hw,
                             w + (size_t)p * in_sz + q * in_hw);
            y[(size_t)i * out + p] = s;
        }
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < out; ++p) y[(size_t)i * out + p] += 1.0f * b[p];
}

/* bcnn_backward_fullc_layer_cpu, src/layers/bcnn_fc_layer.c:183-226:
 * gb += sum_n dy; gW += dy^T x; dx += dy W (all accumulate). */
void orc_fc_backward(const float *x, const float *w, const float *dy, float *gw,
                     float *gb, float *dx, int n, int in_sz, int out) {
    for (int i = 0; i < n; ++i)
        for (int p = 0; p < out; ++p) gb[p] += dy[(size_t)i * out + p];
    for (int p = 0; p < out; ++p)
        for (int q = 0; q < in_sz; ++q) {
            float s = 0.f;
            for (int i = 0; i < n; ++i)
                s += dy[(size_t)i * out + p] * x[(size_t)i * in_sz + q];
            gw[(size_t)p * in_sz + q] += s;
        }
    if (!dx) return;
    for (int i = 0; i < n; ++i)
        for (int q = 0; q < in_sz; ++q) {
            float s = 0.f;
            for (int p = 0; p < out; ++p)
                s += dy[(size_t)i * out + p] * w[(size_t)p * in_sz + q];
            dx[(size_t)i * in_sz + q] += s;
        }
}

/* bcnn_forward_softmax_layer_cpu, src/layers/bcnn_softmax_layer.c:88-155:
 * log-sum-exp form, double exp/log rounded to float at each step. */
void orc_softmax_forward(const float *x, float *y, int n, int c, int hw) {
    for (int b = 0; b < n; ++b)
        for (int i = 0; i < hw; ++i) {
            const float *p = x + (size_t)b * c * hw + i;
            float *q = y + (size_t)b * c * hw + i;
            float vmax = -FLT_MAX, sum = 0.f;
            for (int j = 0; j < c; ++j)
                if (p[(size_t)j * hw] > vmax) vmax = p[(size_t)j * hw];
            for (int j = 0; j < c; ++j)
                sum += (float)exp(p[(size_t)j * hw] - vmax);
            if (sum)
                sum = vmax + (float)log(sum);
            else
                sum = vmax - 100.0f;
            for (int j = 0; j < c; ++j)
                q[(size_t)j * hw] = (float)exp(p[(size_t)j * hw] - sum);
        }
}

/* Correct-batch residual add (the reference's equal-shape path only adds sample
 * 0, src/layers/bcnn_eltwise_layer.c:119-121 -- documented deviation). */
void orc_eltwise_add(const float *a, const float *b, float *y, int sz) {
    for (int i = 0; i < sz; ++i) y[i] = a[i] + b[i];
}

/* ------------------------------------------------------------------ */
/* Concat / upsample (YOLO second head glue)                            */
/* ------------------------------------------------------------------ */

/* bcnn_forward_concat_layer_cpu, src/layers/bcnn_concat_layer.c:107-121: image j of the source
 * (src_sz floats) is copied to dst + dst_offset + j * dst_sz. */
void orc_concat_forward(const float *src, float *dst, int n, int src_sz, int dst_sz, int dst_offset) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < src_sz; ++i)
            dst[(size_t)dst_offset + (size_t)j * dst_sz + i] = src[(size_t)j * src_sz + i];
}

/* bcnn_backward_concat_layer_cpu, :123-142: src_grad += slice of dst_grad (bcnn_axpy, alpha 1). */
void orc_concat_backward(const float *dst_grad, float *src_grad, int n, int src_sz, int dst_sz,
                         int dst_offset) {
    for (int j = 0; j < n; ++j)
        for (int i = 0; i < src_sz; ++i)
            src_grad[(size_t)j * src_sz + i] += dst_grad[(size_t)dst_offset + (size_t)j * dst_sz + i];
}

/* bcnn_forward_upsample_layer_cpu, src/layers/bcnn_upsample_layer.c:86-109. */
void orc_upsample_forward(const float *x, float *y, int n, int c, int h, int w, int size) {
    for (int b = 0; b < n; ++b)
        for (int k = 0; k < c; ++k)
            for (int j = 0; j < h * size; ++j)
                for (int i = 0; i < w * size; ++i) {
                    size_t src_id = (size_t)b * w * h * c + (size_t)k * w * h + (size_t)(j / size) * w + i / size;
                    size_t dst_id = (size_t)b * w * h * c * size * size + (size_t)k * w * h * size * size +
                                    (size_t)j * w * size + i;
                    y[dst_id] = x[src_id];
                }
}

/* bcnn_backward_upsample_layer_cpu, :119-142: the same walk, dx[src] += dy[dst] one at a time. */
void orc_upsample_backward(const float *dy, float *dx, int n, int c, int h, int w, int size) {
    for (int b = 0; b < n; ++b)
        for (int k = 0; k < c; ++k)
            for (int j = 0; j < h * size; ++j)
                for (int i = 0; i < w * size; ++i) {
                    size_t src_id = (size_t)b * w * h * c + (size_t)k * w * h + (size_t)(j / size) * w + i / size;
                    size_t dst_id = (size_t)b * w * h * c * size * size + (size_t)k * w * h * size * size +
                                    (size_t)j * w * size + i;
                    dx[src_id] += dy[dst_id];
                }
}
