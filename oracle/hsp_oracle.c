/*
 * hsp_oracle.c — TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * Plain-C, scalar, single-precision restatement of the HS-Pose hot path
 * (reference: network/fs_net_repo/gcn3d.py of Lynne-Zheng-Linfang/HS-Pose).
 * Each function cites the reference lines it follows.  Pinned against golden
 * vectors produced by the real reference (tests/golden/make_golden.py); see
 * tests/test_oracle_golden.py.  Compiled with -ffp-contract=off so that the
 * only fused multiply-adds are the explicit fmaf() calls.
 *
 * Layouts: row-major contiguous, fp32, indices int32 unless noted.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define DIST_NEIGHBOR 0
#define DIST_NEAREST 1

static inline uint32_t float_orderable(float d) {
  d = d + 0.0f; /* -0 -> +0 */
  uint32_t u;
  memcpy(&u, &d, 4);
  return u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
}

/* |p|^2 as torch.sum(v ** 2, dim=2) evaluates it for short rows: squares are
 * rounded, then added left to right (gcn3d.py:20, :32-33). */
static float sqnorm(const float* p, int D) {
  float q = p[0] * p[0];
  for (int d = 1; d < D; ++d) q = q + p[d] * p[d];
  return q;
}
/* inner product as the K-loop of torch.bmm evaluates it on CPU for tiny K
 * (gcn3d.py:19, :31): a0*b0, then one fused multiply-add per further term. */
static float inner(const float* a, const float* b, int D) {
  float t = a[0] * b[0];
  for (int d = 1; d < D; ++d) t = fmaf(a[d], b[d], t);
  return t;
}

/* get_neighbor_index (gcn3d.py:15-24): formula NEIGHBOR, drop = 1.
 * get_nearest_index  (gcn3d.py:27-36): formula NEAREST, k = 1, drop = 0.
 * Selects the k+drop smallest (distance, index) pairs in ascending order and
 * discards the first `drop` — topk(k+1, largest=False)[..., 1:].  idx (B,M,k). */
void orc_knn(const float* query, const float* cand, int B, int M, int N, int D, int k,
             int drop, int formula, int64_t* idx) {
  const int K = k + drop;
#pragma omp parallel
  {
    uint64_t* best = (uint64_t*)malloc(sizeof(uint64_t) * (size_t)(K + 1));
    float* cq = (float*)malloc(sizeof(float) * (size_t)N);
#pragma omp for schedule(static)
    for (int b = 0; b < B; ++b) {
      const float* cb = cand + (size_t)b * N * D;
      for (int j = 0; j < N; ++j) cq[j] = sqnorm(cb + (size_t)j * D, D);
      for (int i = 0; i < M; ++i) {
        const float* qp = query + ((size_t)b * M + i) * D;
        const float qq = sqnorm(qp, D);
        int n = 0;
        for (int j = 0; j < N; ++j) {
          float t = inner(qp, cb + (size_t)j * D, D);
          float dist;
          if (formula == DIST_NEIGHBOR) dist = ((t * -2.0f) + cq[j]) + qq;
          else dist = (cq[j] + qq) - (2.0f * t);
          uint64_t key = ((uint64_t)float_orderable(dist) << 32) | (uint32_t)j;
          if (n == K && key >= best[K - 1]) continue;
          int p = n < K ? n : K - 1;
          while (p > 0 && best[p - 1] > key) { best[p] = best[p - 1]; --p; }
          best[p] = key;
          if (n < K) ++n;
        }
        for (int r = 0; r < k; ++r)
          idx[((size_t)b * M + i) * k + r] = (int64_t)(best[r + drop] & 0xffffffffu);
      }
    }
    free(best);
    free(cq);
  }
}

/* get_neighbor_direction_norm (gcn3d.py:49-59): F.normalize(nbr - centre). */
static void unit_dir(const float* xyz_b, int i, int j, float* r) {
  float rx = xyz_b[3 * j] - xyz_b[3 * i];
  float ry = xyz_b[3 * j + 1] - xyz_b[3 * i + 1];
  float rz = xyz_b[3 * j + 2] - xyz_b[3 * i + 2];
  float nrm = sqrtf(rx * rx + ry * ry + rz * rz);
  float den = nrm > 1e-12f ? nrm : 1e-12f;
  r[0] = rx / den; r[1] = ry / den; r[2] = rz / den;
}

void orc_direction_norm(const float* xyz, const int32_t* idx, int B, int N, int k, float* out) {
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < N; ++i)
      for (int n = 0; n < k; ++n)
        unit_dir(xyz + (size_t)b * N * 3, i, idx[((size_t)b * N + i) * k + n],
                 out + (((size_t)b * N + i) * k + n) * 3);
}

/* HSlayer_surface.graph_conv (gcn3d.py:92-107); dirn is already
 * F.normalize(directions, dim=0), layout (3, S*C), column j = s*C + c. */
void orc_surface_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn, int B, int N,
                          int k, int S, int C, float* out) {
  const int SC = S * C;
#pragma omp parallel for schedule(static)
  for (int bi = 0; bi < B * N; ++bi) {
    const int b = bi / N, i = bi % N;
    const float* xb = xyz + (size_t)b * N * 3;
    float* o = out + (size_t)bi * C;
    float* mx = (float*)malloc(sizeof(float) * (size_t)SC);
    for (int j = 0; j < SC; ++j) mx[j] = -INFINITY;
    for (int n = 0; n < k; ++n) {
      float r[3];
      unit_dir(xb, i, idx[(size_t)bi * k + n], r);
      for (int j = 0; j < SC; ++j) {
        float th = fmaf(r[2], dirn[2 * SC + j], fmaf(r[1], dirn[SC + j], r[0] * dirn[j]));
        th = th > 0.0f ? th : 0.0f;
        if (th > mx[j]) mx[j] = th;
      }
    }
    for (int c = 0; c < C; ++c) {
      float s = 0.0f;
      for (int t = 0; t < S; ++t) s += mx[t * C + c];
      o[c] = s / (float)S;
    }
    free(mx);
  }
}

/* HS_layer.graph_conv (gcn3d.py:158-181).  P (B,N,(S+1)*C) = fm @ W + bias:
 * centre = P[..., :C], support = P[..., C:].  argmax (B,N,S*C) optional. */
void orc_graph_conv_fwd(const float* xyz, const int32_t* idx, const float* dirn, const float* P,
                        int B, int N, int k, int S, int C, float* out, uint8_t* argmax) {
  const int SC = S * C, LD = (S + 1) * C;
#pragma omp parallel for schedule(static)
  for (int bi = 0; bi < B * N; ++bi) {
    const int b = bi / N, i = bi % N;
    const float* xb = xyz + (size_t)b * N * 3;
    const float* Pb = P + (size_t)b * N * LD;
    float* mx = (float*)malloc(sizeof(float) * (size_t)SC);
    uint8_t* am = (uint8_t*)malloc((size_t)SC);
    for (int j = 0; j < SC; ++j) { mx[j] = -INFINITY; am[j] = 0; }
    for (int n = 0; n < k; ++n) {
      float r[3];
      const int nb = idx[(size_t)bi * k + n];
      unit_dir(xb, i, nb, r);
      const float* sup = Pb + (size_t)nb * LD + C;
      for (int j = 0; j < SC; ++j) {
        float th = fmaf(r[2], dirn[2 * SC + j], fmaf(r[1], dirn[SC + j], r[0] * dirn[j]));
        th = th > 0.0f ? th : 0.0f;
        float a = th * sup[j];
        if (a > mx[j]) { mx[j] = a; am[j] = (uint8_t)n; }
      }
    }
    for (int c = 0; c < C; ++c) {
      float s = 0.0f;
      for (int t = 0; t < S; ++t) s += mx[t * C + c];
      out[(size_t)bi * C + c] = Pb[(size_t)i * LD + c] + s / (float)S;
    }
    if (argmax) memcpy(argmax + (size_t)bi * SC, am, (size_t)SC);
    free(mx);
    free(am);
  }
}

/* Row gather + max over the first kuse neighbours at selected rows:
 * Pool_layer.forward (gcn3d.py:234-246), max part of get_ORL_global (:213-216). */
void orc_gather_max_fwd(const float* feat, const int32_t* idx, const int32_t* rows, int B, int N,
                        int C, int R, int kuse, int kstride, float* out) {
  for (int b = 0; b < B; ++b)
    for (int r = 0; r < R; ++r) {
      const int i = rows ? rows[r] : r;
      for (int c = 0; c < C; ++c) {
        float m = -INFINITY;
        for (int n = 0; n < kuse; ++n) {
          float v = feat[((size_t)b * N + idx[((size_t)b * N + i) * kstride + n]) * C + c];
          if (v > m) m = v;
        }
        out[((size_t)b * R + r) * C + c] = m;
      }
    }
}

/* get_ORL_global (gcn3d.py:211-218): G[b,c] = mean_i max_n feat[b, idx[b,i,n], c]. */
void orc_orl_global_fwd(const float* feat, const int32_t* idx, int B, int N, int C, int k,
                        float* G) {
  for (int b = 0; b < B; ++b)
    for (int c = 0; c < C; ++c) {
      double s = 0.0;
      for (int i = 0; i < N; ++i) {
        float m = -INFINITY;
        for (int n = 0; n < k; ++n) {
          float v = feat[((size_t)b * N + idx[((size_t)b * N + i) * k + n]) * C + c];
          if (v > m) m = v;
        }
        s += (double)m;
      }
      G[(size_t)b * C + c] = (float)(s / (double)N);
    }
}

/* Nearest up-sampling (FaceRecon.py:100-104): out[b,i,col0:col0+C] = feat[b, nn[b,i], :]. */
void orc_upsample_rows_fwd(const float* feat, const int32_t* nn, int B, int Nsrc, int M, int C,
                           float* out, int ldo, int col0) {
  for (int b = 0; b < B; ++b)
    for (int i = 0; i < M; ++i)
      memcpy(out + ((size_t)b * M + i) * ldo + col0,
             feat + ((size_t)b * Nsrc + nn[(size_t)b * M + i]) * C, sizeof(float) * (size_t)C);
}

/* Chamfer forward (tools/pyTorchChamferDistance/chamfer_distance.cu:6-137 /
 * chamfer_distance.cpp:59-87): squared distance to and index of the nearest
 * point of `b` for each point of `a`, direct (x-y)^2 sums, first minimum wins. */
void orc_chamfer_nn(const float* a, const float* b, int B, int N, int M, float* dist,
                    int32_t* idx) {
  for (int o = 0; o < B; ++o)
    for (int i = 0; i < N; ++i) {
      const float* p = a + ((size_t)o * N + i) * 3;
      float best = INFINITY;
      int bj = 0;
      for (int j = 0; j < M; ++j) {
        const float* q = b + ((size_t)o * M + j) * 3;
        float dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
        float d = dx * dx + dy * dy + dz * dz;
        if (d < best) { best = d; bj = j; }
      }
      dist[(size_t)o * N + i] = best;
      idx[(size_t)o * N + i] = bj;
    }
}
