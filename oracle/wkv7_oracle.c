/*
 * oracle/wkv7_oracle.c — CPU restatement of the reference WKV7 recurrence.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under visualrwkv_b200/ may import, link or
 * execute this file; it is the checker for tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py.
 *
 * Follows (never copies) the reference:
 *   forward : VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:10-52   (forward_kernel)
 *             VisualRWKV-v6/v6.xx/RWKV-v7_simple.py:20-32   (fp64 spec loop)
 *   backward: VisualRWKV-v7/v7.00/cuda/wkv7_cuda.cu:54-130  (backward_kernel)
 *
 * Conventions: per (batch b, head h) a state S[i][j],
 * i = value index, j = key index, S_0 = 0.  Arguments in KERNEL order
 * (w, q, k, v, a, b) == python (w, r, k, v, -kk, kk*a).  Tensors are
 * [B,T,H,N] row-major, N = 64 in the product but any N <= 256 works here.
 *
 * Two instantiations of the same source (REAL = double / float):
 *   *_f64 : ground truth (decay via exp(-exp(w)) in double).
 *   *_f32 : emulates the reference kernel's arithmetic: fp32 state, sequential
 *           j-order sums, decay expf(-expf(w)).
 * Inputs are float arrays whose values the caller has already rounded to bf16
 * (the reference kernel reads bf16); outputs are NOT rounded to bf16 here —
 * the Python wrapper does that when it wants the reference's bf16 stores.
 *
 * State checkpoints: s[b][h][t/CH][j][i] = S_ij after step t, (t+1)%CH==0
 * (stored transposed, wkv7_cuda.cu:44-50), CH = 16 in the reference.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define MAXN 256

#define DEFINE_WKV7(REAL, SUF, EXPF)                                                              \
void wkv7_fwd_##SUF(int B, int T, int H, int N, int CH, const float *w_, const float *q_,         \
                    const float *k_, const float *v_, const float *a_, const float *b_,           \
                    REAL *y_, REAL *s_, REAL *sa_, const REAL *s0, REAL *sT) {                    \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                      \
    for (int bb = 0; bb < B; bb++)                                                                 \
        for (int hh = 0; hh < H; hh++) {                                                           \
            REAL *S = (REAL *)calloc((size_t)N * N, sizeof(REAL));                                 \
            REAL wd[MAXN];                                                                         \
            if (s0) memcpy(S, s0 + ((size_t)bb * H + hh) * N * N, sizeof(REAL) * N * N);           \
            for (int t = 0; t < T; t++) {                                                          \
                size_t ind = (((size_t)bb * T + t) * H + hh) * N;                                  \
                for (int j = 0; j < N; j++) wd[j] = EXPF(-EXPF((REAL)w_[ind + j]));                \
                for (int i = 0; i < N; i++) {                                                      \
                    REAL *Si = S + (size_t)i * N;                                                  \
                    REAL sa = 0;                                                                   \
                    for (int j = 0; j < N; j++) sa += (REAL)a_[ind + j] * Si[j];                   \
                    if (sa_) sa_[ind + i] = sa;                                                    \
                    REAL vi = (REAL)v_[ind + i], y = 0;                                            \
                    for (int j = 0; j < N; j++) {                                                  \
                        REAL s = Si[j] * wd[j] + sa * (REAL)b_[ind + j] + (REAL)k_[ind + j] * vi;  \
                        Si[j] = s;                                                                 \
                        y += s * (REAL)q_[ind + j];                                                \
                    }                                                                              \
                    y_[ind + i] = y;                                                               \
                }                                                                                  \
                if (s_ && (t + 1) % CH == 0) {                                                     \
                    REAL *dst = s_ + (((size_t)bb * H + hh) * (T / CH) + t / CH) * N * N;          \
                    for (int i = 0; i < N; i++)                                                    \
                        for (int j = 0; j < N; j++) dst[(size_t)j * N + i] = S[(size_t)i * N + j]; \
                }                                                                                  \
            }                                                                                      \
            if (sT) memcpy(sT + ((size_t)bb * H + hh) * N * N, S, sizeof(REAL) * N * N);           \
            free(S);                                                                               \
        }                                                                                          \
}                                                                                                  \
                                                                                                   \
/* Reverse-time backward exactly as the reference does it: reload the transposed checkpoint at    \
 * chunk ends, un-step the state by division, keep dS in both orientations. */                    \
void wkv7_bwd_##SUF(int B, int T, int H, int N, int CH, const float *w_, const float *q_,         \
                    const float *k_, const float *v_, const float *a_, const float *b_,           \
                    const float *dy_, const REAL *s_, const REAL *sa_, REAL *dw_, REAL *dq_,      \
                    REAL *dk_, REAL *dv_, REAL *da_, REAL *db_) {                                  \
    _Pragma("omp parallel for collapse(2) schedule(dynamic)")                                      \
    for (int bb = 0; bb < B; bb++)                                                                 \
        for (int hh = 0; hh < H; hh++) {                                                           \
            /* St[i][j] = S_{j,i} (column i of S), dS[i][j] = dS_{i,j} */                          \
            REAL *St = (REAL *)calloc((size_t)N * N, sizeof(REAL));                                \
            REAL *dS = (REAL *)calloc((size_t)N * N, sizeof(REAL));                                \
            REAL wd[MAXN], wfac[MAXN], dSb[MAXN];                                                  \
            for (int t = T - 1; t >= 0; t--) {                                                     \
                size_t ind = (((size_t)bb * T + t) * H + hh) * N;                                  \
                const float *q = q_ + ind, *k = k_ + ind, *v = v_ + ind, *a = a_ + ind,            \
                            *b = b_ + ind, *dy = dy_ + ind;                                        \
                const REAL *sa = sa_ + ind;                                                        \
                for (int j = 0; j < N; j++) {                                                      \
                    wfac[j] = -EXPF((REAL)w_[ind + j]);                                            \
                    wd[j] = EXPF(wfac[j]);                                                         \
                }                                                                                  \
                if ((t + 1) % CH == 0) {                                                           \
                    const REAL *src = s_ + (((size_t)bb * H + hh) * (T / CH) + t / CH) * N * N;    \
                    memcpy(St, src, sizeof(REAL) * N * N); /* src[i*N+j] = S_{j,i} */              \
                }                                                                                  \
                for (int i = 0; i < N; i++) {                                                      \
                    REAL *Sti = St + (size_t)i * N;                                                \
                    REAL dq = 0;                                                                   \
                    for (int j = 0; j < N; j++) dq += Sti[j] * (REAL)dy[j];                        \
                    dq_[ind + i] = dq;                                                             \
                    REAL iwi = (REAL)1 / wd[i], ki = k[i], bi = b[i];                              \
                    for (int j = 0; j < N; j++)                                                    \
                        Sti[j] = (Sti[j] - ki * (REAL)v[j] - bi * sa[j]) * iwi;                    \
                }                                                                                  \
                for (int i = 0; i < N; i++)                                                        \
                    for (int j = 0; j < N; j++) dS[(size_t)i * N + j] += (REAL)dy[i] * (REAL)q[j]; \
                for (int i = 0; i < N; i++) {                                                      \
                    REAL dw = 0, dk = 0, dv = 0, db = 0, dsb = 0;                                  \
                    for (int j = 0; j < N; j++) {                                                  \
                        REAL dSji = dS[(size_t)j * N + i], dSij = dS[(size_t)i * N + j];           \
                        dw += dSji * St[(size_t)i * N + j];                                        \
                        dk += dSji * (REAL)v[j];                                                   \
                        dv += dSij * (REAL)k[j];                                                   \
                        dsb += dSij * (REAL)b[j];                                                  \
                        db += dSji * sa[j];                                                        \
                    }                                                                              \
                    dw_[ind + i] = dw * wd[i] * wfac[i];                                           \
                    dk_[ind + i] = dk;                                                             \
                    dv_[ind + i] = dv;                                                             \
                    db_[ind + i] = db;                                                             \
                    dSb[i] = dsb;                                                                  \
                }                                                                                  \
                for (int i = 0; i < N; i++) {                                                      \
                    REAL da = 0;                                                                   \
                    for (int j = 0; j < N; j++) da += St[(size_t)i * N + j] * dSb[j];              \
                    da_[ind + i] = da;                                                             \
                }                                                                                  \
                for (int i = 0; i < N; i++)                                                        \
                    for (int j = 0; j < N; j++)                                                    \
                        dS[(size_t)i * N + j] = dS[(size_t)i * N + j] * wd[j] + dSb[i] * (REAL)a[j]; \
            }                                                                                      \
            free(St);                                                                              \
            free(dS);                                                                              \
        }                                                                                          \
}

DEFINE_WKV7(double, f64, exp)
DEFINE_WKV7(float, f32, expf)

/* Exact adjoint in double that never divides by the decay: recomputes the forward from S_0 = 0,
 * stores every state, then runs the plain reverse-mode recurrence.  O(T*N*N) memory per head, so
 * for small T only; pins the reverse-time reconstruction above. */
void wkv7_bwd_exact_f64(int B, int T, int H, int N, const float *w_, const float *q_,
                        const float *k_, const float *v_, const float *a_, const float *b_,
                        const float *dy_, double *dw_, double *dq_, double *dk_, double *dv_,
                        double *da_, double *db_) {
#pragma omp parallel for collapse(2) schedule(dynamic)
    for (int bb = 0; bb < B; bb++)
        for (int hh = 0; hh < H; hh++) {
            size_t NN = (size_t)N * N;
            double *Sall = (double *)calloc((size_t)(T + 1) * NN, sizeof(double));
            double *SA = (double *)calloc((size_t)T * N, sizeof(double));
            double *dS = (double *)calloc(NN, sizeof(double));
            double wd[MAXN], wfac[MAXN], dsa[MAXN];
            for (int t = 0; t < T; t++) {
                size_t ind = (((size_t)bb * T + t) * H + hh) * N;
                const double *Sp = Sall + (size_t)t * NN;
                double *Sn = Sall + (size_t)(t + 1) * NN;
                for (int j = 0; j < N; j++) wd[j] = exp(-exp((double)w_[ind + j]));
                for (int i = 0; i < N; i++) {
                    double sa = 0;
                    for (int j = 0; j < N; j++) sa += (double)a_[ind + j] * Sp[i * N + j];
                    SA[(size_t)t * N + i] = sa;
                    for (int j = 0; j < N; j++)
                        Sn[i * N + j] = Sp[i * N + j] * wd[j] + sa * (double)b_[ind + j] +
                                        (double)k_[ind + j] * (double)v_[ind + i];
                }
            }
            for (int t = T - 1; t >= 0; t--) {
                size_t ind = (((size_t)bb * T + t) * H + hh) * N;
                const double *Sp = Sall + (size_t)t * NN, *Sn = Sall + (size_t)(t + 1) * NN;
                const double *sa = SA + (size_t)t * N;
                for (int j = 0; j < N; j++) {
                    wfac[j] = -exp((double)w_[ind + j]);
                    wd[j] = exp(wfac[j]);
                }
                /* y = Sn q */
                for (int j = 0; j < N; j++) {
                    double dq = 0;
                    for (int i = 0; i < N; i++) dq += Sn[i * N + j] * (double)dy_[ind + i];
                    dq_[ind + j] = dq;
                }
                for (int i = 0; i < N; i++)
                    for (int j = 0; j < N; j++) dS[i * N + j] += (double)dy_[ind + i] * (double)q_[ind + j];
                /* Sn = Sp*wd + sa b^T + v k^T,  sa = Sp a */
                for (int j = 0; j < N; j++) {
                    double dw = 0, dk = 0, db = 0;
                    for (int i = 0; i < N; i++) {
                        dw += dS[i * N + j] * Sp[i * N + j];
                        dk += dS[i * N + j] * (double)v_[ind + i];
                        db += dS[i * N + j] * sa[i];
                    }
                    dw_[ind + j] = dw * wd[j] * wfac[j];
                    dk_[ind + j] = dk;
                    db_[ind + j] = db;
                }
                for (int i = 0; i < N; i++) {
                    double dv = 0, d = 0;
                    for (int j = 0; j < N; j++) {
                        dv += dS[i * N + j] * (double)k_[ind + j];
                        d += dS[i * N + j] * (double)b_[ind + j];
                    }
                    dv_[ind + i] = dv;
                    dsa[i] = d;
                }
                for (int j = 0; j < N; j++) {
                    double da = 0;
                    for (int i = 0; i < N; i++) da += Sp[i * N + j] * dsa[i];
                    da_[ind + j] = da;
                }
                for (int i = 0; i < N; i++)
                    for (int j = 0; j < N; j++)
                        dS[i * N + j] = dS[i * N + j] * wd[j] + dsa[i] * (double)a_[ind + j];
            }
            free(Sall);
            free(SA);
            free(dS);
        }
}
