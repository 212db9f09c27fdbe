/*
 * oracle/wmd_oracle.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED.
 *
 * Compiled-C restatement of gensim 3.8.x KeyedVectors.wmdistance as the
 * reference calls it (/root/reference/src/wmd.py:31-32 and
 * /root/reference/evaluate/auto/content_preserve.py:43-50), steps S1..S5 of
 * SURVEY.md section 8(c), feeding emd_hat.c (step S6).  It exists so that the
 * GPU parity tests can check 10^4..10^5 pairs in seconds; it is itself checked
 * bit-for-bit against the per-pair numpy path in oracle/wmd_oracle.py
 * (tests/test_oracle_wmd.py), which uses numpy's own float32 kernels exactly
 * the way gensim does.
 *
 * Tokens are embedding-table row numbers (-1 = out of vocabulary).  `rank`
 * gives each row's position in gensim's Dictionary order (Python string sort
 * of the token); NULL means rows are already stored in that order.
 *
 * Build: see oracle/Makefile (-O2 -ffp-contract=off: every float32 op must be
 * separately rounded, as in numpy).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

double emd_hat_gd_metric_double(const double *d1, const double *d2, const double *D, int N,
                                double extra_mass_penalty);

/* numpy FLOAT_pairwise_sum applied to (a-b)**2; each step a rounded float32 op.
 * n < 8: sequential from 0; 8 <= n <= 128: eight strided accumulators combined
 * ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the n%8 tail; n > 128: split at
 * n/2 - (n/2)%8 and recurse.  (SURVEY.md section 7 "Float32 cost fidelity".) */
static float sqdiff_pairwise_f32(const float *a, const float *b, long n)
{
    if (n < 8) {
        float res = 0.f;
        for (long i = 0; i < n; ++i) { float t = a[i] - b[i]; t = t * t; res += t; }
        return res;
    } else if (n <= 128) {
        float r[8];
        for (int k = 0; k < 8; ++k) { float t = a[k] - b[k]; r[k] = t * t; }
        long i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int k = 0; k < 8; ++k) { float t = a[i + k] - b[i + k]; t = t * t; r[k] += t; }
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; ++i) { float t = a[i] - b[i]; t = t * t; res += t; }
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        float lo = sqdiff_pairwise_f32(a, b, n2);
        float hi = sqdiff_pairwise_f32(a + n2, b + n2, n - n2);
        return lo + hi;
    }
}

/* float64(sqrt_f32(sum_f32((x_i - x_j)**2))) -- spec S3 */
float wmd_oracle_dist_f32(const float *a, const float *b, long d)
{
    return sqrtf(sqdiff_pairwise_f32(a, b, d));
}

typedef struct { int row; int key; } rk_t;
static int rk_cmp(const void *x, const void *y)
{
    const rk_t *a = (const rk_t *)x, *b = (const rk_t *)y;
    return (a->key > b->key) - (a->key < b->key);
}

/* unique rows of doc, sorted by rank; returns count, fills rows/counts */
static int uniq_sorted(const int *doc, int n, const int *rank, int *rows, int *counts)
{
    rk_t *t = (rk_t *)malloc(sizeof(rk_t) * (size_t)(n > 0 ? n : 1));
    int m = 0;
    for (int i = 0; i < n; ++i) if (doc[i] >= 0) { t[m].row = doc[i]; t[m].key = rank ? rank[doc[i]] : doc[i]; ++m; }
    qsort(t, (size_t)m, sizeof(rk_t), rk_cmp);
    int u = 0;
    for (int i = 0; i < m; ++i) {
        if (u > 0 && rows[u - 1] == t[i].row) counts[u - 1]++;
        else { rows[u] = t[i].row; counts[u] = 1; ++u; }
    }
    free(t);
    return u;
}

/* status: 0 ok, 1 inf (a side empty after OOV removal), 2 0.0 (one-token union
 * vocabulary), 3 inf (all-zero distance matrix).  SURVEY.md section 8(b). */
double wmd_oracle_pair(const float *table, long d, long stride, const int *rank,
                       const int *doc1, int len1, const int *doc2, int len2, int *status)
{
    int *r1 = (int *)malloc(sizeof(int) * (size_t)(len1 + 1)), *c1 = (int *)malloc(sizeof(int) * (size_t)(len1 + 1));
    int *r2 = (int *)malloc(sizeof(int) * (size_t)(len2 + 1)), *c2 = (int *)malloc(sizeof(int) * (size_t)(len2 + 1));
    int u1 = uniq_sorted(doc1, len1, rank, r1, c1);
    int u2 = uniq_sorted(doc2, len2, rank, r2, c2);
    int n1 = 0, n2 = 0;
    for (int i = 0; i < u1; ++i) n1 += c1[i];
    for (int j = 0; j < u2; ++j) n2 += c2[j];
    double out;
    if (n1 == 0 || n2 == 0) {                                   /* S1 */
        *status = 1; out = INFINITY; goto done;
    }
    {
        /* S2: ids 0..u1-1 = sorted(set(doc1)), then sorted(set(doc2)-set(doc1)) */
        int *id2 = (int *)malloc(sizeof(int) * (size_t)u2);
        int N = u1;
        for (int j = 0; j < u2; ++j) {
            int hit = -1;
            for (int i = 0; i < u1; ++i) if (r1[i] == r2[j]) { hit = i; break; }
            id2[j] = hit >= 0 ? hit : N++;
        }
        if (N == 1) { *status = 2; out = 0.0; free(id2); goto done; }
        double *D  = (double *)calloc((size_t)N * (size_t)N, sizeof(double));
        double *d1 = (double *)calloc((size_t)N, sizeof(double));
        double *d2 = (double *)calloc((size_t)N, sizeof(double));
        double total = 0.0;
        for (int i = 0; i < u1; ++i)                             /* S3 */
            for (int j = 0; j < u2; ++j) {
                int a = i, b = id2[j];
                if (D[a * N + b] != 0.0) continue;
                double v = (double)wmd_oracle_dist_f32(table + (long)r1[i] * stride,
                                                       table + (long)r2[j] * stride, d);
                D[a * N + b] = v; D[b * N + a] = v;
            }
        for (int i = 0; i < N * N; ++i) total += D[i];
        if (total == 0.0) {                                      /* S4 */
            *status = 3; out = INFINITY;
        } else {
            for (int i = 0; i < u1; ++i) d1[i] = (double)c1[i] / (double)n1;        /* S5 */
            for (int j = 0; j < u2; ++j) d2[id2[j]] = (double)c2[j] / (double)n2;
            *status = 0;
            out = emd_hat_gd_metric_double(d1, d2, D, N, -1.0);  /* S6 */
        }
        free(D); free(d1); free(d2); free(id2);
    }
done:
    free(r1); free(c1); free(r2); free(c2);
    return out;
}

/* CSR batch driver; nthreads <= 1 is the single-threaded reference shape
 * (/root/reference/src/main_pretrain.py:120-122: collate runs in the main process).
 * nthreads > 1: pthread workers pulling chunks of 64 pairs from a shared counter. */
typedef struct {
    const float *table; long d, stride; const int *rank;
    const int *ids1; const long long *off1; const int *ids2; const long long *off2;
    long long npairs; double *out; int *status; long long *next;
} batch_job;

static void batch_range(const batch_job *J, long long lo, long long hi)
{
    for (long long p = lo; p < hi; ++p) {
        int st = 0;
        J->out[p] = wmd_oracle_pair(J->table, J->d, J->stride, J->rank,
                                    J->ids1 + J->off1[p], (int)(J->off1[p + 1] - J->off1[p]),
                                    J->ids2 + J->off2[p], (int)(J->off2[p + 1] - J->off2[p]), &st);
        if (J->status) J->status[p] = st;
    }
}

static void *batch_worker(void *arg)
{
    batch_job *J = (batch_job *)arg;
    for (;;) {
        long long lo = __atomic_fetch_add(J->next, 64, __ATOMIC_RELAXED);
        if (lo >= J->npairs) break;
        long long hi = lo + 64 < J->npairs ? lo + 64 : J->npairs;
        batch_range(J, lo, hi);
    }
    return NULL;
}

void wmd_oracle_batch(const float *table, long d, long stride, const int *rank,
                      const int *ids1, const long long *off1,
                      const int *ids2, const long long *off2,
                      long long npairs, double *out, int *status, int nthreads)
{
    long long next = 0;
    batch_job J = { table, d, stride, rank, ids1, off1, ids2, off2, npairs, out, status, &next };
    if (nthreads <= 1) { batch_range(&J, 0, npairs); return; }
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    for (int t = 0; t < nthreads; ++t) pthread_create(&th[t], NULL, batch_worker, &J);
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
}

/* nBOW of one document (spec S5 + SURVEY 8(c).4): unique in-vocabulary rows in
 * canonical (rank) order, int32 counts, FP64 weights count/len. Returns u. */
int wmd_oracle_nbow(const int *rank, const int *doc, int len, int *rows, int *counts, double *weights)
{
    int u = uniq_sorted(doc, len, rank, rows, counts);
    int n = 0;
    for (int i = 0; i < u; ++i) n += counts[i];
    for (int i = 0; i < u; ++i) weights[i] = (double)counts[i] / (double)n;
    return u;
}
