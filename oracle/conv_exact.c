/* ORACLE (test infrastructure only) - never linked into or called by the product.
 *
 * fp32 restatement of the convolution stack of /root/reference/lib/network/rtpose_vgg.py:158-198
 * (nn.Conv2d stride 1 "same" padding + bias (+ReLU), MaxPool2d(2,2)) with a DEFINED accumulation order, so that a
 * second implementation can be compared BIT FOR BIT instead of "within 1e-3":
 *
 *     acc = +0.0f
 *     for tap = (ky, kx) in row-major order:
 *         for c = 0 .. cin-1:  acc = fmaf(in[y+ky-pad][x+kx-pad][c], w[o][c][ky][kx], acc)     (zero outside the image)
 *         if cin % 16 != 0:    acc = fmaf(0, 0, acc)     (the device kernel walks cin in chunks of 16, zero filled)
 *     out = acc + bias[o];  if relu: out = max(out, 0)
 *
 * torch's own CPU convolution (oneDNN) uses an unspecified, ISA-dependent summation order, so it cannot serve as a
 * bit-exact oracle; this restatement is PINNED to it (and so to the reference) through tests/golden/net_368.npz to
 * < 1e-4 max-abs (tests/test_oracle.py).  The product's `fp32` mode (csrc/conv_misc.cu: conv_f32_kernel) follows the
 * same order, which is what makes the end-to-end "identical keypoint assignments" test possible.
 *
 * Layout: activations NHWC fp32 with a channel stride / offset (so concat is a channel slice), weights OIHW.
 */
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#if defined(__AVX2__) && defined(__FMA__)
#include <immintrin.h>
#endif

typedef struct {
    const float *pin, *wp, *bias;
    float* out;
    int N, H, W, PH, PW, cin, cout, cout_p, ks, taps, out_cs, out_off, relu, tail;
    long rows;
    long next;              /* next row to take (atomic) */
} conv_ctx;

static void conv_row(const conv_ctx* k, long r);

static void* conv_worker(void* arg) {
    conv_ctx* k = (conv_ctx*)arg;
    for (;;) {
        const long r = __atomic_fetch_add(&k->next, 1, __ATOMIC_RELAXED);
        if (r >= k->rows) break;
        conv_row(k, r);
    }
    return 0;
}

/* out[n][y][x][out_off + o], in[n][y][x][in_off + c] */
int exact_conv(const float* in, int N, int H, int W, int in_cs, int in_off, int cin, const float* w_oihw,
               const float* bias, int cout, int ks, float* out, int out_cs, int out_off, int relu) {
    const int pad = ks / 2, taps = ks * ks;
    const int PH = H + 2 * pad, PW = W + 2 * pad;
    const int cout_p = (cout + 15) / 16 * 16;
    /* zero-padded copy of the input slice: [N][PH][PW][cin] */
    float* pin = (float*)calloc((size_t)N * PH * PW * cin, sizeof(float));
    /* weights repacked to [tap][cin][cout_p] (zero padded couts are never stored) */
    float* wp = (float*)calloc((size_t)taps * cin * cout_p, sizeof(float));
    if (!pin || !wp) { free(pin); free(wp); return 1; }
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                memcpy(pin + (((size_t)n * PH + y + pad) * PW + x + pad) * cin,
                       in + (((size_t)n * H + y) * W + x) * in_cs + in_off, (size_t)cin * sizeof(float));
    for (int o = 0; o < cout; ++o)
        for (int c = 0; c < cin; ++c)
            for (int t = 0; t < taps; ++t) wp[((size_t)t * cin + c) * cout_p + o] = w_oihw[((size_t)o * cin + c) * taps + t];
    conv_ctx k = {pin, wp, bias, out, N, H, W, PH, PW, cin, cout, cout_p, ks, taps, out_cs, out_off, relu, (cin % 16) != 0,
                  (long)N * H, 0};
    long nt = sysconf(_SC_NPROCESSORS_ONLN);
    const char* env = getenv("ORACLE_THREADS");
    if (env) nt = atol(env);
    if (nt < 1) nt = 1;
    if (nt > 128) nt = 128;
    if (nt > k.rows) nt = k.rows;
    pthread_t th[128];
    long started = 0;
    for (long i = 1; i < nt; ++i)
        if (pthread_create(&th[started], 0, conv_worker, &k) == 0) ++started;
    conv_worker(&k);
    for (long i = 0; i < started; ++i) pthread_join(th[i], 0);
    free(pin);
    free(wp);
    return 0;
}

static void conv_row(const conv_ctx* k, long r) {
    const float *pin = k->pin, *wp = k->wp, *bias = k->bias;
    float* out = k->out;
    const int H = k->H, W = k->W, PH = k->PH, PW = k->PW, cin = k->cin, cout = k->cout, cout_p = k->cout_p, ks = k->ks,
              taps = k->taps, out_cs = k->out_cs, out_off = k->out_off, relu = k->relu, tail = k->tail;
    {
        const int n = (int)(r / H), y = (int)(r % H);
        for (int x0 = 0; x0 < W; x0 += 4) {
            const int np = (W - x0) < 4 ? (W - x0) : 4;
            for (int o0 = 0; o0 < cout_p; o0 += 16) {
#if defined(__AVX2__) && defined(__FMA__)
                __m256 acc[4][2];
                for (int p = 0; p < 4; ++p) acc[p][0] = acc[p][1] = _mm256_setzero_ps();
                for (int t = 0; t < taps; ++t) {
                    const int ky = t / ks, kx = t % ks;
                    const float* src[4];
                    for (int p = 0; p < 4; ++p) {
                        const int xx = (p < np ? x0 + p : x0) + kx;
                        src[p] = pin + (((size_t)n * PH + y + ky) * PW + xx) * cin;
                    }
                    const float* wt = wp + (size_t)t * cin * cout_p + o0;
                    for (int c = 0; c < cin; ++c) {
                        const __m256 w0 = _mm256_loadu_ps(wt + (size_t)c * cout_p), w1 = _mm256_loadu_ps(wt + (size_t)c * cout_p + 8);
                        for (int p = 0; p < 4; ++p) {
                            const __m256 a = _mm256_set1_ps(src[p][c]);
                            acc[p][0] = _mm256_fmadd_ps(a, w0, acc[p][0]);
                            acc[p][1] = _mm256_fmadd_ps(a, w1, acc[p][1]);
                        }
                    }
                    if (tail) {
                        const __m256 z = _mm256_setzero_ps();
                        for (int p = 0; p < 4; ++p) {
                            acc[p][0] = _mm256_fmadd_ps(z, z, acc[p][0]);
                            acc[p][1] = _mm256_fmadd_ps(z, z, acc[p][1]);
                        }
                    }
                }
                for (int p = 0; p < np; ++p) {
                    float v[16];
                    _mm256_storeu_ps(v, acc[p][0]);
                    _mm256_storeu_ps(v + 8, acc[p][1]);
                    float* dst = out + (((size_t)n * H + y) * W + x0 + p) * out_cs + out_off;
                    for (int j = 0; j < 16 && o0 + j < cout; ++j) {
                        volatile float s = v[j] + bias[o0 + j];     /* separate rounding: no contraction */
                        float q = s;
                        if (relu) q = fmaxf(q, 0.f);
                        dst[o0 + j] = q;
                    }
                }
#else
                float acc[4][16];
                for (int p = 0; p < 4; ++p)
                    for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
                for (int t = 0; t < taps; ++t) {
                    const int ky = t / ks, kx = t % ks;
                    const float* wt = wp + (size_t)t * cin * cout_p + o0;
                    for (int c = 0; c < cin; ++c)
                        for (int p = 0; p < np; ++p) {
                            const float a = pin[(((size_t)n * PH + y + ky) * PW + x0 + p + kx) * cin + c];
                            for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(a, wt[(size_t)c * cout_p + j], acc[p][j]);
                        }
                    if (tail)
                        for (int p = 0; p < np; ++p)
                            for (int j = 0; j < 16; ++j) acc[p][j] = fmaf(0.f, 0.f, acc[p][j]);
                }
                for (int p = 0; p < np; ++p) {
                    float* dst = out + (((size_t)n * H + y) * W + x0 + p) * out_cs + out_off;
                    for (int j = 0; j < 16 && o0 + j < cout; ++j) {
                        volatile float s = acc[p][j] + bias[o0 + j];
                        float q = s;
                        if (relu) q = fmaxf(q, 0.f);
                        dst[o0 + j] = q;
                    }
                }
#endif
            }
        }
    }
}

/* MaxPool2d(2, 2), NHWC */
void exact_maxpool(const float* in, int N, int H, int W, int C, float* out) {
    const int Ho = H / 2, Wo = W / 2;
    for (int n = 0; n < N; ++n)
        for (int y = 0; y < Ho; ++y)
            for (int x = 0; x < Wo; ++x)
                for (int c = 0; c < C; ++c) {
                    const float* p = in + (((size_t)n * H + 2 * y) * W + 2 * x) * C + c;
                    out[(((size_t)n * Ho + y) * Wo + x) * C + c] =
                        fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(size_t)W * C], p[(size_t)W * C + C]));
                }
}
