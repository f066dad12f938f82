/* TEST INFRASTRUCTURE ONLY -- plain-C (double precision) restatement of the NBDT tree head and tree-supervision
 * losses, one inner node at a time like the reference (NOT the fused closed form of the CUDA kernels).  Loaded by
 * tests/ and by bench_tree.py's cpu_baseline leg as the checker / CPU baseline; the product path never links it.
 *
 * Parity status: PINNED -- tests/test_oracle_golden.py checks it against oracle/tree_oracle.py (numpy), which is pinned
 * to vectors produced by executing the unmodified reference (tests/golden/make_golden.py).
 *
 * Reference being restated (alvinwan/neural-backed-decision-trees):
 *   node logits = mean of the leaf logits under each child ........ nbdt/model.py:83-99   (get_node_logits)
 *   per-node softmax / outputs ..................................... nbdt/model.py:101-120 (get_all_node_outputs)
 *   class probabilities = product of child probabilities ........... nbdt/model.py:207-242 (Soft...traverse_tree)
 *   L = w_x CE(x, y) + w_t CE(class_probs, y) ...................... nbdt/loss.py:191-203, 260-266 (SoftTreeSupLoss)
 *   hard loss: per node CE on the rows whose target is under it,
 *   grouped by child count, weighted rows/(B N/2) * tsw ............ nbdt/loss.py:212-257 (HardTreeSupLoss), model.py:127-143
 *
 * Hierarchy tables (built by oracle/tree_oracle_c.py from OracleTree): inner node i owns child slots
 * [child_ptr[i], child_ptr[i+1]); slot s lists the class indices under that child in leaf_idx[leaf_ptr[s] .. leaf_ptr[s+1]).
 * Build: oracle/build_oracle.py (gcc -O2 -fopenmp -shared -fPIC). */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
  int C, N, S;
  const int* child_ptr; /* [N+1] */
  const int* leaf_ptr;  /* [S+1] */
  const int* leaf_idx;  /* [leaf_ptr[S]] */
} otree;

int nbdt_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* softmax of z[0..K) into p (double), returns log-sum-exp */
static double softmax_k(const double* z, int K, double* p) {
  double m = z[0], s = 0.0;
  for (int k = 1; k < K; ++k) if (z[k] > m) m = z[k];
  for (int k = 0; k < K; ++k) { p[k] = exp(z[k] - m); s += p[k]; }
  for (int k = 0; k < K; ++k) p[k] /= s;
  return m + log(s);
}

/* per-slot node logits and probabilities of one row (model.py:83-120) */
static void row_nodes(const otree* t, const float* xr, double* nl, double* p) {
  for (int i = 0; i < t->N; ++i) {
    const int b = t->child_ptr[i], e = t->child_ptr[i + 1];
    for (int s = b; s < e; ++s) {
      double acc = 0.0;
      for (int q = t->leaf_ptr[s]; q < t->leaf_ptr[s + 1]; ++q) acc += (double)xr[t->leaf_idx[q]];
      nl[s] = acc / (double)(t->leaf_ptr[s + 1] - t->leaf_ptr[s]);
    }
    softmax_k(nl + b, e - b, p + b);
  }
}

/* class_probs: starts at ones, multiplied per inner node by the probability of the child each class sits under (model.py:207-242) */
static void row_class_probs(const otree* t, const double* p, double* cp) {
  for (int c = 0; c < t->C; ++c) cp[c] = 1.0;
  for (int s = 0; s < t->S; ++s)
    for (int q = t->leaf_ptr[s]; q < t->leaf_ptr[s + 1]; ++q) cp[t->leaf_idx[q]] *= p[s];
}

void nbdt_oracle_soft_probs(const otree* t, const float* x, long long B, long long ld, double* cp_out) {
#pragma omp parallel
  {
    double* nl = (double*)malloc(sizeof(double) * 2 * t->S);
    double* p = nl + t->S;
#pragma omp for schedule(static)
    for (long long b = 0; b < B; ++b) {
      row_nodes(t, x + b * ld, nl, p);
      row_class_probs(t, p, cp_out + b * t->C);
    }
    free(nl);
  }
}

/* CE of one row of "logits" z (double) against target y: value, and softmax(z) - onehot(y) into g */
static double row_ce(const double* z, int C, int y, double* g) {
  const double lse = softmax_k(z, C, g);
  g[y] -= 1.0;
  return lse - z[y];
}

/* SoftTreeSupLoss (loss.py:191-203, 260-266).  Returns the loss; gx (may be NULL) receives dL/dx [B][C] (double). */
double nbdt_oracle_soft_loss(const otree* t, const float* x, const long long* y, long long B, long long ld, double w_xent,
                             double w_tree, double* gx) {
  double loss = 0.0;
#pragma omp parallel reduction(+ : loss)
  {
    const int C = t->C, S = t->S;
    double* buf = (double*)malloc(sizeof(double) * (3 * S + 4 * C));
    double *nl = buf, *p = buf + S, *gp = buf + 2 * S, *cp = buf + 3 * S, *xz = cp + C, *g1 = xz + C, *gcp = g1 + C;
#pragma omp for schedule(static)
    for (long long b = 0; b < B; ++b) {
      const float* xr = x + b * ld;
      const int yy = (int)y[b];
      row_nodes(t, xr, nl, p);
      row_class_probs(t, p, cp);
      for (int c = 0; c < C; ++c) xz[c] = (double)xr[c];
      const double l1 = row_ce(xz, C, yy, g1);   /* criterion(outputs, targets) */
      const double l2 = row_ce(cp, C, yy, gcp);  /* criterion(class PROBABILITIES, targets): loss.py:264-266 */
      loss += (w_xent * l1 + w_tree * l2) / (double)B;
      if (!gx) continue;
      double* gr = gx + b * C;
      for (int c = 0; c < C; ++c) gr[c] = w_xent * g1[c] / (double)B;
      for (int i = 0; i < t->N; ++i) {
        const int s0 = t->child_ptr[i], s1 = t->child_ptr[i + 1];
        double dot = 0.0;
        for (int s = s0; s < s1; ++s) { /* d cp[c] / d p[s] = cp[c] / p[s] for the classes under slot s */
          double acc = 0.0;
          for (int q = t->leaf_ptr[s]; q < t->leaf_ptr[s + 1]; ++q) {
            const int c = t->leaf_idx[q];
            acc += gcp[c] / (double)B * (cp[c] / p[s]);
          }
          gp[s] = acc;
          dot += acc * p[s];
        }
        for (int s = s0; s < s1; ++s) { /* softmax backward, then the backward of the leaf mean */
          const double gnl = p[s] * (gp[s] - dot);
          const double n = (double)(t->leaf_ptr[s + 1] - t->leaf_ptr[s]);
          for (int q = t->leaf_ptr[s]; q < t->leaf_ptr[s + 1]; ++q) gr[t->leaf_idx[q]] += w_tree * gnl / n;
        }
      }
    }
    free(buf);
  }
  return loss;
}

/* HardTreeSupLoss (loss.py:212-257 inside TreeSupLoss.forward, loss.py:191-203).  child_of_class[i*C + c] = slot offset (0..K-1)
 * of the child of inner node i that class c is under, or -1 (model.py:127-143 selects those rows). */
double nbdt_oracle_hard_loss(const otree* t, const int* child_of_class, const float* x, const long long* y, long long B,
                             long long ld, double w_xent, double w_tree, double tsw, double* gx) {
  const int C = t->C, N = t->N;
  const double num_losses = (double)B * (double)N / 2.0; /* loss.py:228 */
  /* rows under every node, summed per child count K (the reference concatenates the nodes of equal K into one CE) */
  int kmax = 0;
  for (int i = 0; i < N; ++i) if (t->child_ptr[i + 1] - t->child_ptr[i] > kmax) kmax = t->child_ptr[i + 1] - t->child_ptr[i];
  double* total_k = (double*)calloc((size_t)kmax + 1, sizeof(double));
  for (long long b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i)
      if (child_of_class[(size_t)i * C + (int)y[b]] >= 0) total_k[t->child_ptr[i + 1] - t->child_ptr[i]] += 1.0;
  double loss = 0.0;
#pragma omp parallel reduction(+ : loss)
  {
    double* buf = (double*)malloc(sizeof(double) * (2 * (size_t)kmax + 2 * (size_t)C));
    double *z = buf, *pz = buf + kmax, *xz = pz + kmax, *g1 = xz + C;
#pragma omp for schedule(static)
    for (long long b = 0; b < B; ++b) {
      const float* xr = x + b * ld;
      const int yy = (int)y[b];
      for (int c = 0; c < C; ++c) xz[c] = (double)xr[c];
      const double l1 = row_ce(xz, C, yy, g1);
      double lt = 0.0;
      double* gr = gx ? gx + b * C : NULL;
      if (gr) for (int c = 0; c < C; ++c) gr[c] = w_xent * g1[c] / (double)B;
      for (int i = 0; i < N; ++i) {
        const int tgt = child_of_class[(size_t)i * C + yy];
        if (tgt < 0) continue;
        const int s0 = t->child_ptr[i], K = t->child_ptr[i + 1] - s0;
        for (int k = 0; k < K; ++k) {
          double acc = 0.0;
          for (int q = t->leaf_ptr[s0 + k]; q < t->leaf_ptr[s0 + k + 1]; ++q) acc += (double)xr[t->leaf_idx[q]];
          z[k] = acc / (double)(t->leaf_ptr[s0 + k + 1] - t->leaf_ptr[s0 + k]);
        }
        const double lse = softmax_k(z, K, pz);
        const double fraction = total_k[K] / num_losses * tsw; /* loss.py:253-255 */
        lt += fraction * (lse - z[tgt]) / total_k[K];          /* CE(mean over the group's rows) */
        if (gr) {
          pz[tgt] -= 1.0;
          for (int k = 0; k < K; ++k) {
            const double gnl = w_tree * fraction / total_k[K] * pz[k];
            const double n = (double)(t->leaf_ptr[s0 + k + 1] - t->leaf_ptr[s0 + k]);
            for (int q = t->leaf_ptr[s0 + k]; q < t->leaf_ptr[s0 + k + 1]; ++q) gr[t->leaf_idx[q]] += gnl / n;
          }
        }
      }
      loss += w_xent * l1 / (double)B + w_tree * lt;
    }
    free(buf);
  }
  free(total_k);
  return loss;
}
