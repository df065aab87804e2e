// K8 — the 19-term loss graph of the 'PoseNet_only' training stage, forward and backward, as three
// kernels.
//
// Replaces, for L1 loss type, the ~2300 element-wise launches per step of
//   losses/fs_net_loss.py:31-76,93-110,123-242   (Rot1, Rot1_cos, Rot2, Rot2_cos, Rot_r_a, Tran, Size, R_con)
//   losses/recon_loss.py:464-649                 (recon_per_p, recon_p_f, recon_point_vote/_r/_t/_s/_self)
//   losses/geometry_loss.py:123-150              (geo_point)
//   losses/prop_loss.py:156-277                  (Prop_pm, Prop_sym_recon, Prop_sym_rt)
// incl. tools/plane_utils.py:24-48 (weighted plane fit — as nine weighted moments per face instead of the
// reference's N x N `diag_embed` weight matrix), tools/rot_utils.py:39-98 and the face normalisation /
// sigmoid of PoseNet9D.py:28-33.
//
//   losses_points_kernel  (CTA = object)   one pass over the N points: every per-object sum the terms need
//                                          (110 floats: residual sums, plane-fit moments, and — for the terms
//                                          that couple a sum over points with predicted pose parameters —
//                                          the value plus its first-order sensitivity sums).
//   losses_object_kernel  (thread = object x input) the per-object algebra.  ONE templated function
//                                          `object_terms<T>` is the single statement of the math; with
//                                          T = float it gives the 19 term contributions, with T = Dual
//                                          (value + one derivative, forward-mode AD) one evaluation per
//                                          input gives the exact gradient w.r.t. the 14 predicted pose
//                                          scalars and the 54 plane-fit moments — no hand-derived
//                                          rotation / 3x3-inverse calculus to get wrong.
//   losses_points_bwd_kernel (thread = point) gradients w.r.t. the raw face head output (N,30) and recon.
// Deterministic (fixed-order tree reductions, no atomics).
#include "common.cuh"

namespace hsp {
namespace loss {

enum Term {
  ROT1, ROT1_COS, ROT2, ROT2_COS, ROT_REG, TRAN, SIZE, RCON, RECON_PER_P, RECON_P_F, VOTE, BB_R, BB_T, BB_S,
  BB_SELF, GEO, PROP_PM, SYM_RECON, SYM_RT, NTERMS
};
enum Wt {
  W_ROT1, W_ROT2, W_ROT_REG, W_TRAN, W_SIZE, W_RCON, W_RECON_N, W_RECON_D, W_RECON_F, W_RECON_V, W_BB_R, W_BB_T,
  W_BB_S, W_BB_SELF, W_GEO_P, W_PROP_PM, W_PROP_SYM, NWT
};
// per-object sums
constexpr int O_SN = 0, O_SD = 6, O_SF = 12, O_MOM = 18, O_GEO = 72, O_PM = 82, O_SYMRECON = 95, O_SYMRT = 96,
              O_VALID = 109, NS = 110;
constexpr int NPRED = 14;   // pg3 pr3 fg fr T3 s3
constexpr int NGT = 23;     // R9 t3 s3 mean_shape3 sym4 obj_id
constexpr int NIN = NPRED + 54;
constexpr int PT_THREADS = 256;

struct Wts { float w[NWT]; };

// ------------------------------------------------------------------ forward-mode dual numbers
struct Dual {
  float v, d;
  __device__ Dual() : v(0.f), d(0.f) {}
  __device__ Dual(float a) : v(a), d(0.f) {}
  __device__ Dual(float a, float b) : v(a), d(b) {}
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return Dual(a.v + b.v, a.d + b.d); }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return Dual(a.v - b.v, a.d - b.d); }
__device__ __forceinline__ Dual operator-(Dual a) { return Dual(-a.v, -a.d); }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return Dual(a.v * b.v, a.d * b.v + a.v * b.d); }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return Dual(q, (a.d - q * b.d) / b.v);
}
__device__ __forceinline__ float val(float a) { return a; }
__device__ __forceinline__ float val(Dual a) { return a.v; }
__device__ __forceinline__ float m_sqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ Dual m_sqrt(Dual a) { const float s = sqrtf(a.v); return Dual(s, a.d / (2.f * s)); }
__device__ __forceinline__ float m_exp(float a) { return expf(a); }
__device__ __forceinline__ Dual m_exp(Dual a) { const float e = expf(a.v); return Dual(e, e * a.d); }
__device__ __forceinline__ float m_sin(float a) { return sinf(a); }
__device__ __forceinline__ Dual m_sin(Dual a) { return Dual(sinf(a.v), cosf(a.v) * a.d); }
__device__ __forceinline__ float m_cos(float a) { return cosf(a); }
__device__ __forceinline__ Dual m_cos(Dual a) { return Dual(cosf(a.v), -sinf(a.v) * a.d); }
__device__ __forceinline__ float m_acos(float a) { return acosf(a); }
__device__ __forceinline__ Dual m_acos(Dual a) { return Dual(acosf(a.v), -a.d / sqrtf(1.f - a.v * a.v)); }
__device__ __forceinline__ float m_abs(float a) { return fabsf(a); }
__device__ __forceinline__ Dual m_abs(Dual a) {   // torch: sign(0) = 0
  return a.v > 0.f ? a : (a.v < 0.f ? -a : Dual(fabsf(a.v), 0.f));
}
__device__ __forceinline__ float m_clamp(float a, float lo, float hi) { return fminf(fmaxf(a, lo), hi); }
__device__ __forceinline__ Dual m_clamp(Dual a, float lo, float hi) {
  return a.v < lo ? Dual(lo, 0.f) : (a.v > hi ? Dual(hi, 0.f) : a);
}

template <typename T> struct V3 { T x, y, z; };
template <typename T> __device__ __forceinline__ V3<T> mk3(T a, T b, T c) { V3<T> r; r.x = a; r.y = b; r.z = c; return r; }
template <typename T> __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return mk3<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> __device__ __forceinline__ V3<T> scale(V3<T> a, T s) { return mk3<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __device__ __forceinline__ V3<T> add(V3<T> a, V3<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __device__ __forceinline__ V3<T> sub(V3<T> a, V3<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __device__ __forceinline__ T norm(V3<T> a) { return m_sqrt(dot(a, a)); }
template <typename T> __device__ __forceinline__ T comp(V3<T> a, int k) { return k == 0 ? a.x : (k == 1 ? a.y : a.z); }
template <typename T> __device__ __forceinline__ V3<T> lift(V3<float> a) { return mk3<T>(T(a.x), T(a.y), T(a.z)); }
template <typename T> __device__ __forceinline__ T mean_abs3(V3<T> a) { return (m_abs(a.x) + m_abs(a.y) + m_abs(a.z)) / T(3.f); }

// F.normalize: v / max(|v|, 1e-12)
template <typename T> __device__ __forceinline__ V3<T> normalize12(V3<T> a) {
  const T n = norm(a);
  return val(n) > 1e-12f ? scale(a, T(1.f) / n) : scale(a, T(1e12f));
}

// tools/rot_utils.py:39-65 with r orthogonal to y and z:  v' = cos(t) v + sin(t) (r x v)
template <typename T>
__device__ void vertical_rot_vec(float c1, float c2, V3<T> y, V3<T> z, V3<T>& new_y, V3<T>& new_z) {
  V3<T> r = cross(y, z);
  r = scale(r, T(1.f) / (norm(r) + T(1e-8f)));
  const T c = m_clamp(dot(y, z), -1.f + 1e-6f, 1.f - 1e-6f);
  const T excess = m_acos(c) - T(1.57079632679489662f);
  const T t1 = T(c2 / (c1 + c2)) * excess, t2 = T(c1 / (c1 + c2)) * excess;
  new_y = add(scale(y, m_cos(t1)), scale(cross(r, y), m_sin(t1)));
  new_z = sub(scale(z, m_cos(t2)), scale(cross(r, z), m_sin(t2)));
}
// tools/rot_utils.py:76-85: columns (x', y', z')
template <typename T>
__device__ void rot_mat_y_first(V3<T> y, V3<T> x, V3<T>& cx, V3<T>& cy, V3<T>& cz) {
  cy = normalize12(y);
  cz = normalize12(cross(x, cy));
  cx = cross(cy, cz);
}

// ------------------------------------------------------------------ per-object constants
struct ObjConst {
  float R[9];        // gt_R row-major: R[3*j + k]
  float t[3], re_s[3], mean_shape[3], gt_s[3];
  int sym0, nosym, obj5, y_refl, yx_refl, no_refl, skip;
  float bs, N, rn, valid;   // rn: renormalisation bs / valid (1 if none valid)
};
__device__ __forceinline__ void load_const(const float* gt, float valid, int B, int N, ObjConst& c) {
#pragma unroll
  for (int i = 0; i < 9; ++i) c.R[i] = gt[i];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    c.t[i] = gt[9 + i];
    c.gt_s[i] = gt[12 + i];
    c.mean_shape[i] = gt[15 + i];
    c.re_s[i] = gt[12 + i] + gt[15 + i];
  }
  const float s0 = gt[18], s1 = gt[19], s2 = gt[20], s3 = gt[21];
  c.sym0 = s0 == 1.f;
  c.nosym = s0 == 0.f;
  c.obj5 = gt[22] == 5.f;
  const bool any_rest = (s1 + s2 + s3) > 0.f;
  c.y_refl = (s0 == 1.f) && any_rest;
  c.yx_refl = (s0 == 0.f) && (s1 == 1.f);
  c.no_refl = (s0 == 0.f) && (s1 != 1.f);
  c.skip = (s0 == 1.f) && !any_rest;
  c.bs = (float)B;
  c.N = (float)N;
  c.valid = valid;
  c.rn = valid > 0.f ? (float)B / valid : 1.f;
}
__device__ __forceinline__ V3<float> axis(const ObjConst& c, int k) { return mk3<float>(c.R[k], c.R[3 + k], c.R[6 + k]); }
__device__ __forceinline__ float axis_mask(const ObjConst& c, int a) {   // recon_loss.py:546-554
  return a == 1 ? 1.f : (a == 0 ? (float)(c.nosym && !c.obj5) : (float)c.nosym);
}

// Predicted rotation of Prop_pm (prop_loss.py:156-187) and mirror normal of Prop_sym_rt (:239-249)
template <typename T>
__device__ void pred_frames(const ObjConst& c, float fg, float fr, V3<T> pg, V3<T> pr, V3<T>& Rx, V3<T>& Ry, V3<T>& Rz,
                            V3<T>& pz) {
  V3<T> ny, nx;
  if (c.sym0) vertical_rot_vec<T>(fg, 1e-5f, pg, lift<T>(axis(c, 0)), ny, nx);
  else vertical_rot_vec<T>(fg, fr, pg, pr, ny, nx);
  rot_mat_y_first(ny, nx, Rx, Ry, Rz);
  pz = cross(pr, pg);
  pz = scale(pz, T(1.f) / (norm(pz) + T(1e-8f)));
}

// ------------------------------------------------------------------ the per-object algebra (single source)
template <typename T>
__device__ void object_terms(const ObjConst& c, const Wts& W, const float* __restrict__ S /*sums*/, V3<T> pg, V3<T> pr,
                             T fg, T fr, V3<T> pt, V3<T> ps, const T* mom /*54*/, const float* pred0 /*14 floats*/,
                             T* out /*NTERMS*/) {
  const float ib = 1.f / c.bs, i3b = 1.f / (3.f * c.bs), ibn = 1.f / (c.bs * c.N), i3bn = 1.f / (3.f * c.bs * c.N);
  const V3<float> gy = axis(c, 1), gx = axis(c, 0);
  const float fg0 = pred0[6], fr0 = pred0[7];
  const V3<float> pg0 = mk3<float>(pred0[0], pred0[1], pred0[2]), pr0 = mk3<float>(pred0[3], pred0[4], pred0[5]);
  const V3<float> pt0 = mk3<float>(pred0[8], pred0[9], pred0[10]);
  const T zero(0.f);
  // ---- fs_net terms
  const V3<T> dg = sub(pg, lift<T>(gy)), dr = sub(pr, lift<T>(gx));
  out[ROT1] = T(W.w[W_ROT1] * i3b) * (m_abs(dg.x) + m_abs(dg.y) + m_abs(dg.z));
  out[ROT1_COS] = T(W.w[W_ROT1] * 2.f * ib) * (T(1.f) - dot(pg, lift<T>(gy)));
  out[ROT2] = c.nosym ? T(W.w[W_ROT2] * c.rn * i3b) * (m_abs(dr.x) + m_abs(dr.y) + m_abs(dr.z)) : zero;
  out[ROT2_COS] = c.nosym ? T(W.w[W_ROT2] * c.rn * 2.f * ib) * (T(1.f) - dot(pr, lift<T>(gx))) : zero;
  out[ROT_REG] = c.nosym ? T(W.w[W_ROT_REG] * c.rn * ib) * m_abs(dot(pg, pr)) : zero;
  const V3<T> dt = sub(pt, lift<T>(mk3<float>(c.t[0], c.t[1], c.t[2])));
  const V3<T> ds = sub(ps, lift<T>(mk3<float>(c.gt_s[0], c.gt_s[1], c.gt_s[2])));
  out[TRAN] = T(W.w[W_TRAN] * i3b) * (m_abs(dt.x) + m_abs(dt.y) + m_abs(dt.z));
  out[SIZE] = T(W.w[W_SIZE] * i3b) * (m_abs(ds.x) + m_abs(ds.y) + m_abs(ds.z));
  {
    T r = m_abs(m_exp(T(-13.7f) * dot(dg, dg)) - fg);
    if (c.nosym) r = r + m_abs(m_exp(T(-13.7f) * dot(dr, dr)) - fr);
    out[RCON] = T(W.w[W_RCON] * ib) * r;
  }
  // ---- recon_6face per-point terms: linear in the per-object sums (their point gradients are local)
  {
    const float in6 = 1.f / (6.f * c.bs * c.N);
    float rn = S[O_SN + 1] + S[O_SN + 4] + (c.nosym ? S[O_SN] + S[O_SN + 2] + S[O_SN + 3] + S[O_SN + 5] : 0.f);
    float rd = 0.f, rf = 0.f;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      rd += axis_mask(c, f % 3) * S[O_SD + f];
      rf += axis_mask(c, f % 3) * S[O_SF + f];
    }
    out[RECON_PER_P] = T((W.w[W_RECON_N] * rn + W.w[W_RECON_D] * rd) * in6);
    out[RECON_P_F] = T(W.w[W_RECON_F] * rf * in6);
  }
  // ---- recon_6face voting terms: weighted plane fit per face from its nine moments
  {
    V3<T> fn[6];
    T fc[6];
    T vote = zero;
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const T* m = mom + 9 * f;   // sw sx sy sxx sxy syy sxz syz sz
      const T sw = m[0], sx = m[1], sy = m[2], sxx = m[3], sxy = m[4], syy = m[5];
      // inverse of [[sxx,sxy,sx],[sxy,syy,sy],[sx,sy,sw]] (adjugate / determinant), times (sxz, syz, sz)
      const T A = syy * sw - sy * sy, Bc = -(sxy * sw - sy * sx), Cc = sxy * sy - syy * sx;
      const T det = sxx * A + sxy * Bc + sx * Cc;
      const T i00 = A, i01 = -(sxy * sw - sx * sy), i02 = sxy * sy - sx * syy;
      const T i10 = Bc, i11 = sxx * sw - sx * sx, i12 = -(sxx * sy - sx * sxy);
      const T i20 = Cc, i21 = -(sxx * sy - sxy * sx), i22 = sxx * syy - sxy * sxy;
      const T X0 = (i00 * m[6] + i01 * m[7] + i02 * m[8]) / det;
      const T X1 = (i10 * m[6] + i11 * m[7] + i12 * m[8]) / det;
      const T X2 = (i20 * m[6] + i21 * m[7] + i22 * m[8]) / det;
      const T dn_norm = X0 * X0 + X1 * X1 + T(1.f);
      const V3<T> dn = scale(mk3<T>(X0 * X2, X1 * X2, -X2), T(1.f) / (dn_norm + T(1e-8f)));
      V3<T> n = scale(dn, T(1.f) / norm(dn));
      T off = X2 / m_sqrt(dn_norm);
      const int a = f % 3;
      const float sg = f < 3 ? 1.f : -1.f;
      const V3<float> ax = scale(axis(c, a), sg);
      if (val(dot(n, lift<T>(ax))) < 0.f) { n = scale(n, T(-1.f)); off = -off; }
      fn[f] = n;
      fc[f] = off;
      const float half = c.re_s[a] * 0.5f;
      const V3<float> corner = mk3<float>(c.t[0] + ax.x * half, c.t[1] + ax.y * half, c.t[2] + ax.z * half);
      const V3<float> dn_gt = scale(ax, -dot(ax, corner));
      vote = vote + T(axis_mask(c, a)) * mean_abs3(sub(dn, lift<T>(dn_gt)));
    }
    V3<T> new_y, new_x;
    vertical_rot_vec<T>(fg0, fr0, pg, pr, new_y, new_x);
    const V3<T> new_z = cross(new_x, new_y);
    const V3<T> pa[3] = {new_x, new_y, new_z};
    T geo_r = zero, geo_t = zero, geo_s = zero, self_cal = zero;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const T mk(axis_mask(c, a));
      geo_r = geo_r + mk * (mean_abs3(sub(fn[a], pa[a])) + mean_abs3(add(fn[a + 3], pa[a])));
      const T dis_up = m_abs(dot(fn[a], pt) + fc[a]), dis_down = m_abs(dot(fn[a + 3], pt) + fc[a + 3]);
      geo_t = geo_t + mk * m_abs(dis_down - dis_up);
      const T hs = (comp(ps, a) + T(c.mean_shape[a])) * T(0.5f);
      geo_s = geo_s + mk * (m_abs(hs - dis_up) + m_abs(hs - dis_down));
      self_cal = self_cal + mk * mean_abs3(add(fn[a], fn[a + 3]));
      if (a != 1) self_cal = self_cal + mk * (m_abs(dot(fn[1], fn[a])) + m_abs(dot(fn[4], fn[a + 3])));
    }
    const float i6 = 1.f / (6.f * c.bs);
    out[VOTE] = T(W.w[W_RECON_V] * i6) * vote;
    out[BB_R] = T(W.w[W_BB_R] * i6) * geo_r;
    out[BB_T] = T(W.w[W_BB_T] * i6) * geo_t;
    out[BB_S] = T(W.w[W_BB_S] * i6) * geo_s;
    out[BB_SELF] = T(W.w[W_BB_SELF] * i6) * self_cal;
  }
  // ---- terms that couple a sum over points with predicted pose parameters: value + first-order
  // sensitivities were accumulated by the point pass at (pg0, pr0, pt0, ...); the derivative part of the
  // (theta - theta0) factors below carries the chain rule
  const V3<T> dpt = sub(pt, lift<T>(pt0));
  {
    const float* G = S + O_GEO;   // Vy Sy3 cy Vx Sx3 cx
    const V3<float> sy = mk3<float>(G[1] - G[4] * pt0.x, G[2] - G[4] * pt0.y, G[3] - G[4] * pt0.z);
    T gy_term = T(G[0]) + dot(lift<T>(sy), sub(pg, lift<T>(pg0))) + dot(lift<T>(scale(pg0, -G[4])), dpt);
    T gx_term = zero;
    if (c.nosym) {
      const V3<float> sx = mk3<float>(G[6] - G[9] * pt0.x, G[7] - G[9] * pt0.y, G[8] - G[9] * pt0.z);
      gx_term = T(c.rn) * (T(G[5]) + dot(lift<T>(sx), sub(pr, lift<T>(pr0))) + dot(lift<T>(scale(pr0, -G[9])), dpt));
    }
    out[GEO] = T(W.w[W_GEO_P] * ibn) * (gy_term + gx_term);
  }
  V3<T> Rc[3], pz;
  pred_frames<T>(c, fg0, fr0, pg, pr, Rc[0], Rc[1], Rc[2], pz);
  V3<float> Rc0[3], pz0;
  pred_frames<float>(c, fg0, fr0, pg0, pr0, Rc0[0], Rc0[1], Rc0[2], pz0);
  {
    const float* P = S + O_PM;    // V, S_k (3 vectors), c_k (3)
    T acc = T(P[0]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float ck = P[10 + k];
      const V3<float> sk = mk3<float>(P[1 + 3 * k] - ck * pt0.x, P[2 + 3 * k] - ck * pt0.y, P[3 + 3 * k] - ck * pt0.z);
      acc = acc + dot(lift<T>(sk), sub(Rc[k], lift<T>(Rc0[k]))) + dot(lift<T>(scale(Rc0[k], -ck)), dpt);
    }
    out[PROP_PM] = T(W.w[W_PROP_PM] * i3bn) * acc;
  }
  out[SYM_RECON] = T(W.w[W_PROP_SYM] * i3bn * S[O_SYMRECON]);
  {
    const float* Q = S + O_SYMRT;   // V, sigma (3), M (9) = sum sigma rel^T
    const V3<float> sig = mk3<float>(Q[1], Q[2], Q[3]);
    T acc = T(Q[0]);
    auto Mv = [&](V3<float> v) {   // (M + M^T) v
      return mk3<float>((Q[4] + Q[4]) * v.x + (Q[5] + Q[7]) * v.y + (Q[6] + Q[10]) * v.z,
                        (Q[7] + Q[5]) * v.x + (Q[8] + Q[8]) * v.y + (Q[9] + Q[11]) * v.z,
                        (Q[10] + Q[6]) * v.x + (Q[11] + Q[9]) * v.y + (Q[12] + Q[12]) * v.z);
    };
    if (c.y_refl) {
      const V3<float> dVdpt = sub(scale(sig, 2.f), scale(pg0, 2.f * dot(pg0, sig)));
      const V3<float> dVdg = scale(Mv(pg0), 2.f);
      acc = acc + dot(lift<T>(dVdpt), dpt) + dot(lift<T>(dVdg), sub(pg, lift<T>(pg0)));
    } else if (c.yx_refl) {
      const V3<float> dVdpz = scale(Mv(pz0), -2.f);
      const V3<float> dVdpt = scale(pz0, 2.f * dot(pz0, sig));
      acc = acc + dot(lift<T>(dVdpz), sub(pz, lift<T>(pz0))) + dot(lift<T>(dVdpt), dpt);
    }
    out[SYM_RT] = T(W.w[W_PROP_SYM] * i3bn) * acc;
  }
}

// ------------------------------------------------------------------ per-point data (one face-head row)
struct PointFaces {
  float n[6][3], rn[6] /* 1/|v| */, d[6], c[6];   // loss order: x+ y+ z+ x- y- z-
};
__device__ __forceinline__ void load_faces(const float* __restrict__ row /*30*/, PointFaces& F) {
  const int perm[6] = {1, 0, 2, 3, 5, 4};
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int s = perm[f];
    const float x = row[3 * s], y = row[3 * s + 1], z = row[3 * s + 2];
    const float inv = 1.f / sqrtf(x * x + y * y + z * z);
    F.n[f][0] = x * inv; F.n[f][1] = y * inv; F.n[f][2] = z * inv;
    F.rn[f] = inv;
    F.d[f] = row[18 + s];
    F.c[f] = 1.f / (1.f + expf(-row[24 + s]));
  }
}
__device__ __forceinline__ float sgn(float a) { return a > 0.f ? 1.f : (a < 0.f ? -1.f : 0.f); }

struct PointGeom {       // what both point passes need of one point
  float p[3], proj[3], rel[3];
  float target[3], pcb[3];   // Prop_sym targets (0 when the branch is off)
};
__device__ __forceinline__ void point_geom(const ObjConst& c, const float* pc, const float* pt, const float* pg,
                                           const float* pz, PointGeom& g) {
#pragma unroll
  for (int j = 0; j < 3; ++j) { g.p[j] = pc[j]; g.rel[j] = pc[j] - pt[j]; }
#pragma unroll
  for (int k = 0; k < 3; ++k)
    g.proj[k] = (g.p[0] - c.t[0]) * c.R[k] + (g.p[1] - c.t[1]) * c.R[3 + k] + (g.p[2] - c.t[2]) * c.R[6 + k];
  float m[3] = {0.f, 0.f, 0.f};
  if (c.yx_refl) { m[0] = g.proj[0]; m[1] = g.proj[1]; m[2] = -g.proj[2]; }
  if (c.y_refl) { m[0] = -g.proj[0]; m[1] = g.proj[1]; m[2] = -g.proj[2]; }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    g.target[j] = 0.f;
    g.pcb[j] = 0.f;
    if (c.yx_refl || c.y_refl) g.target[j] = c.R[3 * j] * m[0] + c.R[3 * j + 1] * m[1] + c.R[3 * j + 2] * m[2] + c.t[j];
    if (c.no_refl) g.target[j] = g.p[j];
  }
  if (c.y_refl) {
    const float a = g.rel[0] * pg[0] + g.rel[1] * pg[1] + g.rel[2] * pg[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) g.pcb[j] = g.p[j] + 2.f * (a * pg[j] - g.rel[j]);
  } else if (c.yx_refl) {
    const float a = g.rel[0] * pz[0] + g.rel[1] * pz[1] + g.rel[2] * pz[2];
#pragma unroll
    for (int j = 0; j < 3; ++j) g.pcb[j] = g.p[j] - 2.f * a * pz[j];
  }
}

// ------------------------------------------------------------------ kernel 1: sums over the points
__global__ void __launch_bounds__(PT_THREADS)
losses_points_kernel(const float* __restrict__ face, const float* __restrict__ recon, const float* __restrict__ PC,
                     const float* __restrict__ pred, const float* __restrict__ gt, int B, int N,
                     float* __restrict__ sums) {
  __shared__ float s_red[PT_THREADS / 32][NS];
  __shared__ float s_valid;
  __shared__ float s_frame[12];   // p_R columns (9) + p_z (3)
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    float v = 0.f;
    for (int i = 0; i < B; ++i) v += gt[(size_t)i * NGT + 18] == 0.f ? 1.f : 0.f;
    s_valid = v;
  }
  __syncthreads();
  ObjConst c;
  load_const(gt + (size_t)b * NGT, s_valid, B, N, c);
  const float* pr0 = pred + (size_t)b * NPRED;
  const float pg[3] = {pr0[0], pr0[1], pr0[2]}, pred_r[3] = {pr0[3], pr0[4], pr0[5]};
  const float pt[3] = {pr0[8], pr0[9], pr0[10]};
  if (tid == 0) {
    V3<float> Rx, Ry, Rz, pz;
    pred_frames<float>(c, pr0[6], pr0[7], mk3<float>(pg[0], pg[1], pg[2]), mk3<float>(pred_r[0], pred_r[1], pred_r[2]), Rx, Ry,
                       Rz, pz);
    const V3<float> cols[3] = {Rx, Ry, Rz};
    for (int k = 0; k < 3; ++k) { s_frame[3 * k] = cols[k].x; s_frame[3 * k + 1] = cols[k].y; s_frame[3 * k + 2] = cols[k].z; }
    s_frame[9] = pz.x; s_frame[10] = pz.y; s_frame[11] = pz.z;
  }
  __syncthreads();
  float Rc[3][3], pz[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { Rc[k][0] = s_frame[3 * k]; Rc[k][1] = s_frame[3 * k + 1]; Rc[k][2] = s_frame[3 * k + 2]; }
  pz[0] = s_frame[9]; pz[1] = s_frame[10]; pz[2] = s_frame[11];

  float acc[NS - 1];
#pragma unroll
  for (int i = 0; i < NS - 1; ++i) acc[i] = 0.f;
  for (int n = tid; n < N; n += PT_THREADS) {
    const size_t pi = (size_t)b * N + n;
    PointFaces F;
    load_faces(face + pi * 30, F);
    PointGeom g;
    point_geom(c, PC + pi * 3, pt, pg, pz, g);
#pragma unroll
    for (int f = 0; f < 6; ++f) {
      const int a = f % 3;
      const float sg = f < 3 ? 1.f : -1.f;
      const float ax[3] = {sg * c.R[a], sg * c.R[3 + a], sg * c.R[6 + a]};
      const float dgt = c.re_s[a] * 0.5f - sg * g.proj[a];
      acc[O_SN + f] += 1.f - (F.n[f][0] * ax[0] + F.n[f][1] * ax[1] + F.n[f][2] * ax[2]);
      acc[O_SD + f] += fabsf(F.d[f] - dgt);
      float cc2 = 0.f;
#pragma unroll
      for (int j = 0; j < 3; ++j) { const float m = F.n[f][j] * F.d[f] - ax[j] * dgt; cc2 += m * m; }
      acc[O_SF + f] += fabsf(expf(-303.5f * cc2) - F.c[f]);
      const float qx = g.p[0] + F.d[f] * F.n[f][0], qy = g.p[1] + F.d[f] * F.n[f][1], qz = g.p[2] + F.d[f] * F.n[f][2];
      const float w = F.c[f];
      float* m = acc + O_MOM + 9 * f;
      m[0] += w; m[1] += w * qx; m[2] += w * qy; m[3] += w * qx * qx; m[4] += w * qx * qy; m[5] += w * qy * qy;
      m[6] += w * qx * qz; m[7] += w * qy * qz; m[8] += w * qz;
    }
    {   // geo_point
      const float ry = g.rel[0] * pg[0] + g.rel[1] * pg[1] + g.rel[2] * pg[2] - g.proj[1];
      const float s = sgn(ry);
      float* G = acc + O_GEO;
      G[0] += fabsf(ry); G[1] += s * g.p[0]; G[2] += s * g.p[1]; G[3] += s * g.p[2]; G[4] += s;
      if (c.nosym) {
        const float rx = g.rel[0] * pred_r[0] + g.rel[1] * pred_r[1] + g.rel[2] * pred_r[2] - g.proj[0];
        const float sx = sgn(rx);
        G[5] += fabsf(rx); G[6] += sx * g.p[0]; G[7] += sx * g.p[1]; G[8] += sx * g.p[2]; G[9] += sx;
      }
    }
    {   // Prop_pm
      float* P = acc + O_PM;
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const float u = g.rel[0] * Rc[k][0] + g.rel[1] * Rc[k][1] + g.rel[2] * Rc[k][2] - g.proj[k];
        const float s = sgn(u);
        P[0] += fabsf(u);
        P[1 + 3 * k] += s * g.p[0]; P[2 + 3 * k] += s * g.p[1]; P[3 + 3 * k] += s * g.p[2];
        P[10 + k] += s;
      }
    }
    if (!c.skip) {   // Prop_sym
      const float* re = recon + pi * 3;
      float* Q = acc + O_SYMRT;
#pragma unroll
      for (int j = 0; j < 3; ++j) acc[O_SYMRECON] += fabsf(g.target[j] - re[j]);
      if (c.y_refl || c.yx_refl) {
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const float df = g.pcb[j] - re[j], s = sgn(df);
          Q[0] += fabsf(df);
          Q[1 + j] += s;
          Q[4 + 3 * j] += s * g.rel[0]; Q[5 + 3 * j] += s * g.rel[1]; Q[6 + 3 * j] += s * g.rel[2];
        }
      }
    }
  }
  // fixed-order reduction: lanes (shuffle tree), then the warps in order
#pragma unroll
  for (int i = 0; i < NS - 1; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) s_red[warp][i] = v;
  }
  __syncthreads();
  for (int i = tid; i < NS - 1; i += PT_THREADS) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < PT_THREADS / 32; ++w) v += s_red[w][i];
    sums[(size_t)b * NS + i] = v;
  }
  if (tid == 0) sums[(size_t)b * NS + O_VALID] = s_valid;
}

// ------------------------------------------------------------------ kernel 2: per-object algebra
// mode 0: thread = object, term contributions (B, NTERMS).  mode 1: thread = (object, input i < 68):
// d(sum_k g_k term_k)/d input_i  ->  gpred (B,14) | gmom (B,54).
__global__ void __launch_bounds__(128)
losses_object_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ sums,
                     Wts W, int B, int N, int mode, const float* __restrict__ gterm, float* __restrict__ pieces,
                     float* __restrict__ gpred, float* __restrict__ gmom) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = mode == 0 ? t : t / NIN, i = mode == 0 ? -1 : t % NIN;
  if (b >= B) return;
  const float* S = sums + (size_t)b * NS;
  ObjConst c;
  load_const(gt + (size_t)b * NGT, S[O_VALID], B, N, c);
  const float* p0 = pred + (size_t)b * NPRED;
  if (mode == 0) {
    float out[NTERMS];
    object_terms<float>(c, W, S, mk3<float>(p0[0], p0[1], p0[2]), mk3<float>(p0[3], p0[4], p0[5]), p0[6], p0[7],
                        mk3<float>(p0[8], p0[9], p0[10]), mk3<float>(p0[11], p0[12], p0[13]), S + O_MOM, p0, out);
#pragma unroll
    for (int k = 0; k < NTERMS; ++k) pieces[(size_t)b * NTERMS + k] = out[k];
    return;
  }
  Dual in[NPRED], mom[54], out[NTERMS];
#pragma unroll
  for (int j = 0; j < NPRED; ++j) in[j] = Dual(p0[j], j == i ? 1.f : 0.f);
  for (int j = 0; j < 54; ++j) mom[j] = Dual(S[O_MOM + j], (NPRED + j) == i ? 1.f : 0.f);
  object_terms<Dual>(c, W, S, mk3<Dual>(in[0], in[1], in[2]), mk3<Dual>(in[3], in[4], in[5]), in[6], in[7],
                     mk3<Dual>(in[8], in[9], in[10]), mk3<Dual>(in[11], in[12], in[13]), mom, p0, out);
  float g = 0.f;
#pragma unroll
  for (int k = 0; k < NTERMS; ++k) g += gterm[k] * out[k].d;
  if (i < NPRED) gpred[(size_t)b * NPRED + i] = g;
  else gmom[(size_t)b * 54 + (i - NPRED)] = g;
}

// ------------------------------------------------------------------ kernel 3: gradients at the points
__global__ void __launch_bounds__(PT_THREADS)
losses_points_bwd_kernel(const float* __restrict__ face, const float* __restrict__ recon, const float* __restrict__ PC,
                         const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ sums,
                         const float* __restrict__ gmom, const float* __restrict__ gterm, Wts W, int B, int N,
                         float* __restrict__ gface, float* __restrict__ grecon) {
  __shared__ float s_frame[3];
  __shared__ float s_gm[54];
  const int b = blockIdx.y, tid = threadIdx.x;
  ObjConst c;
  load_const(gt + (size_t)b * NGT, sums[(size_t)b * NS + O_VALID], B, N, c);
  const float* pr0 = pred + (size_t)b * NPRED;
  const float pg[3] = {pr0[0], pr0[1], pr0[2]}, pt[3] = {pr0[8], pr0[9], pr0[10]};
  if (tid == 0) {
    V3<float> Rx, Ry, Rz, pz;
    pred_frames<float>(c, pr0[6], pr0[7], mk3<float>(pg[0], pg[1], pg[2]), mk3<float>(pr0[3], pr0[4], pr0[5]), Rx, Ry, Rz, pz);
    s_frame[0] = pz.x; s_frame[1] = pz.y; s_frame[2] = pz.z;
  }
  if (tid < 54) s_gm[tid] = gmom[(size_t)b * 54 + tid];
  __syncthreads();
  const int n = blockIdx.x * PT_THREADS + tid;
  if (n >= N) return;
  const float pz[3] = {s_frame[0], s_frame[1], s_frame[2]};
  const size_t pi = (size_t)b * N + n;
  PointFaces F;
  load_faces(face + pi * 30, F);
  PointGeom g;
  point_geom(c, PC + pi * 3, pt, pg, pz, g);
  const float in6 = 1.f / (6.f * c.bs * c.N);
  const float kN = gterm[RECON_PER_P] * W.w[W_RECON_N] * in6, kD = gterm[RECON_PER_P] * W.w[W_RECON_D] * in6;
  const float kF = gterm[RECON_P_F] * W.w[W_RECON_F] * in6;
  const int perm[6] = {1, 0, 2, 3, 5, 4};
  float* go = gface + pi * 30;
#pragma unroll
  for (int f = 0; f < 6; ++f) {
    const int a = f % 3;
    const float sg = f < 3 ? 1.f : -1.f;
    const float ax[3] = {sg * c.R[a], sg * c.R[3 + a], sg * c.R[6 + a]};
    const float dgt = c.re_s[a] * 0.5f - sg * g.proj[a];
    const float mN = a == 1 ? 1.f : (float)c.nosym, mD = axis_mask(c, a);
    float gn[3], gd, gc;
    // normals / distances / confidences of the per-point terms
#pragma unroll
    for (int j = 0; j < 3; ++j) gn[j] = -kN * mN * ax[j];
    gd = kD * mD * sgn(F.d[f] - dgt);
    float m[3], cc2 = 0.f, mn = 0.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) { m[j] = F.n[f][j] * F.d[f] - ax[j] * dgt; cc2 += m[j] * m[j]; mn += m[j] * F.n[f][j]; }
    const float e = expf(-303.5f * cc2);
    const float sF = kF * mD * sgn(e - F.c[f]);
    gc = -sF;
    const float de = sF * (-303.5f) * e * 2.f;
#pragma unroll
    for (int j = 0; j < 3; ++j) gn[j] += de * F.d[f] * m[j];
    gd += de * mn;
    // plane-fit moments (weights = detached confidences)
    {
      const float* G = s_gm + 9 * f;
      const float qx = g.p[0] + F.d[f] * F.n[f][0], qy = g.p[1] + F.d[f] * F.n[f][1], qz = g.p[2] + F.d[f] * F.n[f][2];
      const float w = F.c[f];
      const float dq[3] = {w * (G[1] + 2.f * qx * G[3] + qy * G[4] + qz * G[6]),
                           w * (G[2] + qx * G[4] + 2.f * qy * G[5] + qz * G[7]),
                           w * (G[8] + qx * G[6] + qy * G[7])};
#pragma unroll
      for (int j = 0; j < 3; ++j) { gd += dq[j] * F.n[f][j]; gn[j] += F.d[f] * dq[j]; }
    }
    // back through n = v / |v|, c = sigmoid(logit)
    const float ng = F.n[f][0] * gn[0] + F.n[f][1] * gn[1] + F.n[f][2] * gn[2];
    const int s = perm[f];
#pragma unroll
    for (int j = 0; j < 3; ++j) go[3 * s + j] = (gn[j] - F.n[f][j] * ng) * F.rn[f];
    go[18 + s] = gd;
    go[24 + s] = gc * F.c[f] * (1.f - F.c[f]);
  }
  // recon (Prop_sym): d/d recon of |target - recon| and |PC_b - recon|
  const float i3bn = 1.f / (3.f * c.bs * c.N);
  const float kR = gterm[SYM_RECON] * W.w[W_PROP_SYM] * i3bn, kT = gterm[SYM_RT] * W.w[W_PROP_SYM] * i3bn;
  const float* re = recon + pi * 3;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    float gr = 0.f;
    if (!c.skip) {
      gr -= kR * sgn(g.target[j] - re[j]);
      if (c.y_refl || c.yx_refl) gr -= kT * sgn(g.pcb[j] - re[j]);
    }
    grecon[pi * 3 + j] = gr;
  }
}

}  // namespace loss
}  // namespace hsp

extern "C" int hsp_losses_num_terms(void) { return hsp::loss::NTERMS; }
extern "C" int hsp_losses_num_sums(void) { return hsp::loss::NS; }

extern "C" int hsp_losses_fwd(const float* face, const float* recon, const float* PC, const float* pred,
                              const float* gt, const float* weights, int B, int N, float* sums, float* pieces,
                              void* stream) {
  using namespace hsp;
  using namespace hsp::loss;
  if (!face || !recon || !PC || !pred || !gt || !weights || !sums || !pieces || B <= 0 || N <= 0 || B > 65535)
    return HSP_EINVAL;
  Wts W;
  for (int i = 0; i < NWT; ++i) W.w[i] = weights[i];   // HOST array of the 17 FLAGS weights
  cudaStream_t st = (cudaStream_t)stream;
  losses_points_kernel<<<B, PT_THREADS, 0, st>>>(face, recon, PC, pred, gt, B, N, sums);
  HSP_LAUNCH_CHECK();
  losses_object_kernel<<<(B + 127) / 128, 128, 0, st>>>(pred, gt, sums, W, B, N, 0, nullptr, pieces, nullptr, nullptr);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}

extern "C" int hsp_losses_bwd(const float* face, const float* recon, const float* PC, const float* pred,
                              const float* gt, const float* weights, const float* sums, const float* gterm, int B,
                              int N, float* gface, float* grecon, float* gpred, float* gmom_ws, void* stream) {
  using namespace hsp;
  using namespace hsp::loss;
  if (!face || !recon || !PC || !pred || !gt || !weights || !sums || !gterm || !gface || !grecon || !gpred ||
      !gmom_ws || B <= 0 || N <= 0 || B > 65535)
    return HSP_EINVAL;
  Wts W;
  for (int i = 0; i < NWT; ++i) W.w[i] = weights[i];
  cudaStream_t st = (cudaStream_t)stream;
  const int threads = B * NIN;
  losses_object_kernel<<<(threads + 127) / 128, 128, 0, st>>>(pred, gt, sums, W, B, N, 1, gterm, nullptr, gpred, gmom_ws);
  HSP_LAUNCH_CHECK();
  losses_points_bwd_kernel<<<dim3((N + PT_THREADS - 1) / PT_THREADS, B), PT_THREADS, 0, st>>>(
      face, recon, PC, pred, gt, sums, gmom_ws, gterm, W, B, N, gface, grecon);
  HSP_LAUNCH_CHECK();
  return HSP_OK;
}
