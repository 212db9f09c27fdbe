/*
 * oracle/emd_hat.c -- TEST INFRASTRUCTURE ONLY (CPU oracle). PARITY UNPINNED.
 *
 * Plain-C restatement of the arithmetic the reference reaches through
 *     /root/reference/src/wmd.py:32                         (self.model.wv.wmdistance)
 *     /root/reference/evaluate/auto/content_preserve.py:47  (wmd_model.wv.wmdistance)
 *     /root/reference/evaluate/auto/transfer_intensity.py:11 (pyemd.emd, direct)
 * i.e. gensim 3.8.x KeyedVectors.wmdistance -> pyemd 0.5.1 emd() ->
 * emd_hat_gd_metric<double> (lib/emd_hat_impl.hpp, lib/min_cost_flow.hpp).
 * Neither gensim nor pyemd is vendored in /root/reference or installed here,
 * so this file restates their *published algorithm* from the normative spec in
 * SURVEY.md section 8(c), steps S1..S6.  "Parity unpinned": the reference has
 * no tests or golden vectors for this path; what pins this file instead is
 *   - pyemd's own known-answer tests (tests/test_oracle_emd.py),
 *   - two independent exact solvers (scipy HiGHS LP, networkx network simplex),
 *   - numpy itself for the float32 pairwise-sum order (tests/test_oracle_wmd.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libwmd_b200.so) never does.
 *
 * Nothing here is copied from pyemd: the graph formulation (threshold node,
 * artificial node, isolated-node pre-flow) follows its published structure so
 * that the integer optimum is the one pyemd would return; the shortest-path
 * machinery is an O(V^2) array Dijkstra written for this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

#define EMD_INF_I64 ((i64)0x3fffffffffffffffLL)

/* ------------------------------------------------------------------------ */
/* Integer min-cost flow by successive shortest paths.                       */
/* Spec S6(e); pyemd lib/min_cost_flow.hpp: repeatedly take the node with    */
/* the largest remaining excess, Dijkstra (reduced costs) to the nearest     */
/* deficit node, push the bottleneck.  The optimum VALUE is unique, so the   */
/* tie-breaking of the original does not matter.                             */
/* ------------------------------------------------------------------------ */
typedef struct {
    int nv, ne;          /* nodes, directed arcs incl. reverse arcs (ne even) */
    int *head;           /* head[v] = first arc of v, -1 = none              */
    int *next;           /* next arc in v's list                             */
    int *to;
    i64 *cost;           /* cost of arc a; arc a^1 is its reverse (-cost)     */
    i64 *cap;            /* residual capacity                                 */
} mcf_graph;

static void mcf_init(mcf_graph *g, int nv, int max_arcs)
{
    g->nv = nv; g->ne = 0;
    g->head = (int *)malloc(sizeof(int) * (size_t)nv);
    g->next = (int *)malloc(sizeof(int) * (size_t)max_arcs * 2);
    g->to   = (int *)malloc(sizeof(int) * (size_t)max_arcs * 2);
    g->cost = (i64 *)malloc(sizeof(i64) * (size_t)max_arcs * 2);
    g->cap  = (i64 *)malloc(sizeof(i64) * (size_t)max_arcs * 2);
    for (int v = 0; v < nv; ++v) g->head[v] = -1;
}
static void mcf_free(mcf_graph *g)
{
    free(g->head); free(g->next); free(g->to); free(g->cost); free(g->cap);
}
static void mcf_add(mcf_graph *g, int u, int v, i64 c)
{
    int a = g->ne;
    g->to[a] = v; g->cost[a] = c;  g->cap[a] = EMD_INF_I64; g->next[a] = g->head[u]; g->head[u] = a;
    g->to[a + 1] = u; g->cost[a + 1] = -c; g->cap[a + 1] = 0; g->next[a + 1] = g->head[v]; g->head[v] = a + 1;
    g->ne += 2;
}

/* e[v] > 0 supply, < 0 demand, sum(e) == 0.  Returns sum(flow*cost). */
static i64 mcf_solve(mcf_graph *g, i64 *e)
{
    const int nv = g->nv;
    i64 *pi   = (i64 *)calloc((size_t)nv, sizeof(i64));
    i64 *dist = (i64 *)malloc(sizeof(i64) * (size_t)nv);
    int *parc = (int *)malloc(sizeof(int) * (size_t)nv);
    char *done = (char *)malloc((size_t)nv);
    i64 total = 0;

    for (;;) {
        int k = -1; i64 best = 0;
        for (int v = 0; v < nv; ++v) if (e[v] > best) { best = e[v]; k = v; }
        if (k < 0) break;

        for (int v = 0; v < nv; ++v) { dist[v] = EMD_INF_I64; done[v] = 0; parc[v] = -1; }
        dist[k] = 0;
        int l = -1;
        for (;;) {
            int u = -1; i64 du = EMD_INF_I64;
            for (int v = 0; v < nv; ++v) if (!done[v] && dist[v] < du) { du = dist[v]; u = v; }
            if (u < 0) break;
            done[u] = 1;
            if (e[u] < 0) { l = u; break; }
            for (int a = g->head[u]; a >= 0; a = g->next[a]) {
                if (g->cap[a] <= 0) continue;
                int w = g->to[a];
                if (done[w]) continue;
                i64 nd = du + g->cost[a] + pi[u] - pi[w];
                if (nd < dist[w]) { dist[w] = nd; parc[w] = a; }
            }
        }
        if (l < 0) { total = -1; break; }             /* infeasible: cannot happen (artificial node) */
        for (int v = 0; v < nv; ++v)
            pi[v] += (done[v] && dist[v] < dist[l]) ? dist[v] : dist[l];

        i64 delta = e[k] < -e[l] ? e[k] : -e[l];
        for (int v = l; v != k; v = g->to[parc[v] ^ 1])
            if (g->cap[parc[v]] < delta) delta = g->cap[parc[v]];
        for (int v = l; v != k; v = g->to[parc[v] ^ 1]) {
            int a = parc[v];
            if (g->cap[a] < EMD_INF_I64 / 2) g->cap[a] -= delta;
            g->cap[a ^ 1] += delta;
            total += delta * g->cost[a];
        }
        e[k] -= delta; e[l] += delta;
    }
    free(pi); free(dist); free(parc); free(done);
    return total;
}

/* ------------------------------------------------------------------------ */
/* emd_hat on integral types.  Spec S6(e); pyemd lib/emd_hat_impl.hpp         */
/* (emd_hat_impl_integral_types): heavier side supplies, threshold node       */
/* absorbs the surplus at zero cost and feeds sinks at cost maxC, arcs with   */
/* cost == maxC are routed through it, nodes touching only the threshold are  */
/* pre-flowed, an artificial node keeps the network connected.                */
/* P,Q: residual masses (after cancellation), C: N*N row-major.               */
/* ------------------------------------------------------------------------ */
i64 emd_hat_integral(const i64 *Pc, const i64 *Qc, const i64 *C, int N, i64 extra_mass_penalty)
{
    i64 sumP = 0, sumQ = 0;
    for (int i = 0; i < N; ++i) { sumP += Pc[i]; sumQ += Qc[i]; }
    const i64 *P = Pc, *Q = Qc;
    i64 diff = sumP - sumQ;
    if (sumQ > sumP) { P = Qc; Q = Pc; diff = sumQ - sumP; }   /* C assumed symmetric, as upstream */

    const int THR = 2 * N, ART = 2 * N + 1, NV = 2 * N + 2;
    i64 *b = (i64 *)calloc((size_t)NV, sizeof(i64));
    for (int i = 0; i < N; ++i) { b[i] = P[i]; b[N + i] = Q[i]; }
    b[THR] = -diff;

    i64 maxC = 0;
    for (int i = 0; i < N * N; ++i) if (C[i] > maxC) maxC = C[i];
    if (extra_mass_penalty == -1) extra_mass_penalty = maxC;

    char *linked = (char *)calloc((size_t)NV, 1);   /* has a regular (non-threshold) arc */
    for (int i = 0; i < N; ++i) {
        if (b[i] == 0) continue;
        for (int j = 0; j < N; ++j) {
            if (b[N + j] == 0 || C[i * N + j] == maxC) continue;
            linked[i] = 1; linked[N + j] = 1;
        }
    }
    for (int i = N; i < 2 * N; ++i) b[i] = -b[i];

    i64 pre_flow_cost = 0;
    int *name = (int *)malloc(sizeof(int) * (size_t)NV);
    int nn = 0;
    for (int i = 0; i < 2 * N; ++i) {
        name[i] = -1;
        if (b[i] == 0) continue;
        if (linked[i]) { name[i] = nn++; }
        else {
            if (i >= N) pre_flow_cost -= b[i] * maxC;   /* isolated sink: fed by the threshold */
            b[THR] += b[i];
        }
    }
    name[THR] = nn++; name[ART] = nn++;

    mcf_graph g;
    mcf_init(&g, nn, N * N + 2 * N + 2 * NV + 8);
    i64 *bb = (i64 *)calloc((size_t)nn, sizeof(i64));
    for (int i = 0; i < NV; ++i) if (name[i] >= 0) bb[name[i]] = b[i];
    for (int i = 0; i < N; ++i) {
        if (name[i] < 0) continue;
        for (int j = 0; j < N; ++j) {
            if (name[N + j] < 0 || C[i * N + j] == maxC) continue;
            mcf_add(&g, name[i], name[N + j], C[i * N + j]);
        }
    }
    for (int i = 0; i < N; ++i)     if (name[i] >= 0)     mcf_add(&g, name[i], name[THR], 0);
    for (int j = 0; j < N; ++j)     if (name[N + j] >= 0) mcf_add(&g, name[THR], name[N + j], maxC);
    for (int i = 0; i < ART; ++i) {
        if (name[i] < 0) continue;
        mcf_add(&g, name[i], name[ART], maxC + 1);
        mcf_add(&g, name[ART], name[i], maxC + 1);
    }
    i64 mcf_dist = mcf_solve(&g, bb);
    mcf_free(&g);
    free(bb); free(name); free(linked); free(b);
    return pre_flow_cost + mcf_dist + diff * extra_mass_penalty;
}

/* ------------------------------------------------------------------------ */
/* emd_hat_gd_metric<double>.  Spec S6(a)-(f): metric pre-flow cancellation,  */
/* x1e6 quantisation to long long, integer solve, un-normalise.               */
/* d1,d2: histograms [N]; D: N*N row-major ground distances.                  */
/* ------------------------------------------------------------------------ */
double emd_hat_gd_metric_double(const double *d1, const double *d2, const double *D, int N,
                                double extra_mass_penalty)
{
    const double MULT_FACTOR = 1000000;
    double *P = (double *)malloc(sizeof(double) * (size_t)N);
    double *Q = (double *)malloc(sizeof(double) * (size_t)N);
    for (int i = 0; i < N; ++i) {                       /* S6(a) */
        P[i] = d1[i]; Q[i] = d2[i];
        if (P[i] < Q[i]) { Q[i] -= P[i]; P[i] = 0; }
        else             { P[i] -= Q[i]; Q[i] = 0; }
    }
    double sumP = 0.0, sumQ = 0.0, maxC = D[0];         /* S6(b): original histograms, id order */
    for (int i = 0; i < N; ++i) {
        sumP += d1[i]; sumQ += d2[i];
        for (int j = 0; j < N; ++j) if (D[i * N + j] > maxC) maxC = D[i * N + j];
    }
    double minSum = sumP < sumQ ? sumP : sumQ;
    double maxSum = sumP < sumQ ? sumQ : sumP;
    double PQn = MULT_FACTOR / maxSum;                  /* S6(c) */
    double Cn  = MULT_FACTOR / maxC;
    i64 *iP = (i64 *)malloc(sizeof(i64) * (size_t)N);
    i64 *iQ = (i64 *)malloc(sizeof(i64) * (size_t)N);
    i64 *iC = (i64 *)malloc(sizeof(i64) * (size_t)N * (size_t)N);
    for (int i = 0; i < N; ++i) {                       /* S6(d) */
        iP[i] = (i64)floor(P[i] * PQn + 0.5);
        iQ[i] = (i64)floor(Q[i] * PQn + 0.5);
        for (int j = 0; j < N; ++j) iC[i * N + j] = (i64)floor(D[i * N + j] * Cn + 0.5);
    }
    double dist = (double)emd_hat_integral(iP, iQ, iC, N, 0);   /* S6(e) */
    dist = dist / PQn;                                  /* S6(f) */
    dist = dist / Cn;
    if (extra_mass_penalty == -1) extra_mass_penalty = maxC;
    dist += (maxSum - minSum) * extra_mass_penalty;
    free(P); free(Q); free(iP); free(iQ); free(iC);
    return dist;
}

/* Exposed so tests can cross-check the integer stage against networkx. */
void emd_hat_quantise(const double *d1, const double *d2, const double *D, int N,
                      i64 *iP, i64 *iQ, i64 *iC)
{
    double sumP = 0.0, sumQ = 0.0, maxC = D[0];
    for (int i = 0; i < N; ++i) {
        sumP += d1[i]; sumQ += d2[i];
        for (int j = 0; j < N; ++j) if (D[i * N + j] > maxC) maxC = D[i * N + j];
    }
    double maxSum = sumP < sumQ ? sumQ : sumP;
    double PQn = 1000000 / maxSum, Cn = 1000000 / maxC;
    for (int i = 0; i < N; ++i) {
        double p = d1[i], q = d2[i];
        if (p < q) { q -= p; p = 0; } else { p -= q; q = 0; }
        iP[i] = (i64)floor(p * PQn + 0.5);
        iQ[i] = (i64)floor(q * PQn + 0.5);
        for (int j = 0; j < N; ++j) iC[i * N + j] = (i64)floor(D[i * N + j] * Cn + 0.5);
    }
}
