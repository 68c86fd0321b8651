// Host build of diffusion_extensions_b200/csrc/so3d_math.cuh -- TEST HARNESS ONLY.
// Lets the CPU-only test tier (no GPU in the build container) exercise the exact per-rotation
// arithmetic the sm_100a kernels inline, against the oracle.  Never loaded by the product.
#define SO3D_HOST_ONLY 1
#include "../../diffusion_extensions_b200/csrc/so3d_math.cuh"

using namespace so3d;

static inline Mat3 ld(const float* p) { Mat3 m; for (int i = 0; i < 9; ++i) m.m[i] = p[i]; return m; }
static inline void st(float* p, const Mat3& m) { for (int i = 0; i < 9; ++i) p[i] = m.m[i]; }

extern "C" {
void hm_axis_angle(const float* R, float* axis, float* ang, long n) {
  for (long i = 0; i < n; ++i) { AxisAngle a = axis_angle(ld(R + 9 * i)); axis[3*i]=a.axis.x; axis[3*i+1]=a.axis.y; axis[3*i+2]=a.axis.z; ang[i]=a.theta; }
}
void hm_log_vec(const float* R, float* v, long n) {
  for (long i = 0; i < n; ++i) { Vec3 a = log_vec(ld(R + 9 * i)); v[3*i]=a.x; v[3*i+1]=a.y; v[3*i+2]=a.z; }
}
void hm_aa_to_rmat(const float* axis, const float* ang, float* R, long n) {
  for (long i = 0; i < n; ++i) st(R + 9 * i, aa_to_rmat(Vec3{axis[3*i], axis[3*i+1], axis[3*i+2]}, ang[i]));
}
void hm_exp_vec(const float* v, float* R, long n) {
  for (long i = 0; i < n; ++i) st(R + 9 * i, exp_vec(Vec3{v[3*i], v[3*i+1], v[3*i+2]}));
}
void hm_scale(const float* R, const float* s, float* out, long n) {
  for (long i = 0; i < n; ++i) st(out + 9 * i, scale_rot(ld(R + 9 * i), s[i]));
}
// what the row-engine kernels run for the L0 maps (lean primitives)
void hm_log_vec_fast(const float* R, float* v, long n) {
  for (long i = 0; i < n; ++i) { Vec3 a = log_vec_fast(ld(R + 9 * i)); v[3*i]=a.x; v[3*i+1]=a.y; v[3*i+2]=a.z; }
}
void hm_exp_vec_fast(const float* v, float* R, long n) {
  for (long i = 0; i < n; ++i) st(R + 9 * i, quat_to_mat_unit(quat_exp_vec(Vec3{v[3*i], v[3*i+1], v[3*i+2]})));
}
void hm_scale_fast(const float* R, const float* s, float* out, long n) {
  for (long i = 0; i < n; ++i) st(out + 9 * i, scale_rot_fast(ld(R + 9 * i), s[i]));
}
void hm_quat_to_rmat(const float* q, float* R, long n) {
  for (long i = 0; i < n; ++i) st(R + 9 * i, quat_to_rmat(q[4*i], q[4*i+1], q[4*i+2], q[4*i+3]));
}
void hm_rmat_to_quat(const float* R, float* q, long n) {
  for (long i = 0; i < n; ++i) rmat_to_quat(ld(R + 9 * i), q + 4 * i);
}
void hm_closed_f32(const float* w, const float* eps, float* logf, float* g, long n) {
  for (long i = 0; i < n; ++i) igso3_closed_f32(w[i], eps[i], logf + i, g + i);
}
void hm_closed_f64(const double* w, const double* eps, double* f, long n, int quirks) {
  for (long i = 0; i < n; ++i) f[i] = igso3_closed_f64(w[i], eps[i], quirks);
}
void hm_sincos_fast(const float* x, float* s, float* c, long n) {
  for (long i = 0; i < n; ++i) sincos_fast(x[i], s + i, c + i);
}
void hm_atan2_pos(const float* y, const float* x, float* r, long n) {
  for (long i = 0; i < n; ++i) r[i] = atan2_pos(y[i], x[i]);
}
void hm_axis_angle_fast(const float* R, float* axis, float* ang, long n) {
  for (long i = 0; i < n; ++i) { AxisAngleF a = axis_angle_fast(ld(R + 9 * i)); axis[3*i]=a.axis.x; axis[3*i+1]=a.axis.y; axis[3*i+2]=a.axis.z; ang[i]=a.theta; }
}
// fused reverse-step mean on quaternions -> matrices (mean and x0_hat)
void hm_p_mean_quat(const float* x, const float* pred, const float* a, const float* b, const float* c1, const float* c2, float* mean,
                    float* x0h, long n) {
  for (long i = 0; i < n; ++i) {
    Quat qh;
    const Quat qm = p_mean_quat(ld(x + 9 * i), Vec3{pred[3*i], pred[3*i+1], pred[3*i+2]}, a[i], b[i], c1[i], c2[i], &qh);
    st(mean + 9 * i, quat_to_mat_unit(qm));
    st(x0h + 9 * i, quat_to_mat_unit(qh));
  }
}
void hm_series(const float* w, const float* eps, float* F, float* Fp, long n, int L) {
  for (long i = 0; i < n; ++i) { SeriesAcc a = igso3_series_terms(w[i], eps[i], L); F[i] = a.F; Fp[i] = a.dF; }
}
void hm_angle_from_uniform(const float* trap, const float* loc, const float* u, float* ang, long n) {
  for (long i = 0; i < n; ++i) ang[i] = igso3_angle_from_uniform(trap, loc, u[i]);
}
// inverse CDF through the 16-byte guide records (built here from the same row) -- must equal the full search
void hm_angle_from_record(const float* trap, const float* loc, const float* u, float* ang, long n) {
  for (long i = 0; i < n; ++i) {
    const GuideRec rec = make_guide_rec(trap, guide_bucket(u[i]));
    ang[i] = igso3_angle_from_record(trap, loc, rec, u[i]);
  }
}
// inverse CDF through the shared-memory guide (uint16 counts at sg_edge(0 .. kSgBuckets), built here like stage_cdf does);
// one_probe[i] = 1 when the bucket of u[i] holds at most one grid point
void hm_angle_from_smem_guide(const float* trap, const float* loc, const float* u, float* ang, int* bucket, int* one_probe, long n) {
  static unsigned short guide[kGuideStride];
  for (int k = 0; k <= kSgBuckets; ++k) guide[k] = (unsigned short)cdf_count_le(trap, sg_edge(k), 0, kCdf);
  for (long i = 0; i < n; ++i) {
    ang[i] = igso3_angle_from_uniform_guided(trap, loc, guide, u[i]);
    bucket[i] = sg_bucket(u[i]);
    one_probe[i] = (guide[bucket[i] + 1] - guide[bucket[i]] <= 1) ? 1 : 0;
  }
}
void hm_sg_edges(float* e) { for (int k = 0; k <= kSgBuckets; ++k) e[k] = sg_edge(k); }
int hm_sg_buckets() { return kSgBuckets; }
// record index of every u, the [a, b] range its record serves, and whether one record resolves the lookup
void hm_guide_bucket(const float* trap, const float* u, int* k, float* a, float* b, int* one_load, long n) {
  for (long i = 0; i < n; ++i) {
    k[i] = guide_bucket(u[i]);
    guide_rec_range(k[i], a + i, b + i);
    const GuideRec rec = make_guide_rec(trap, k[i]);
    const int lo = (int)(rec.lohi & 0xffffu), hi = (int)(rec.lohi >> 16);
    one_load[i] = (hi - lo <= 1 && lo < kCdf - 1) ? 1 : 0;
  }
}
int hm_guide_records() { return kGuideRecs; }
void hm_philox(unsigned long long seed, unsigned long long row0, unsigned long long offset, unsigned* out, long n) {
  for (long i = 0; i < n; ++i) { U4 r = philox4x32_10(seed, row0 + i, offset); out[4*i]=r.x; out[4*i+1]=r.y; out[4*i+2]=r.z; out[4*i+3]=r.w; }
}
// the launcher-side key schedule (what the kernels use) must reproduce the plain generator bit for bit
void hm_philox_keyed(unsigned long long seed, unsigned long long row0, unsigned long long offset, unsigned* out, long n) {
  const PhiloxKey k = make_philox_key(seed, offset);
  for (long i = 0; i < n; ++i) { U4 r = philox4x32_10(k, row0 + i); out[4*i]=r.x; out[4*i+1]=r.y; out[4*i+2]=r.z; out[4*i+3]=r.w; }
}
void hm_draw(unsigned long long seed, unsigned long long row0, unsigned long long offset, float* axis, float* u, long n) {
  for (long i = 0; i < n; ++i) { NoiseDraw d = draw_axis_u(seed, row0 + i, offset); axis[3*i]=d.axis.x; axis[3*i+1]=d.axis.y; axis[3*i+2]=d.axis.z; u[i]=d.u; }
}
void hm_log_bwd(const float* R, const float* G, float* out, long n) {
  for (long i = 0; i < n; ++i) st(out + 9 * i, log_bwd(ld(R + 9 * i), ld(G + 9 * i)));
}
void hm_aa_bwd(const float* axis, const float* ang, const float* G, float* gaxis, float* gang, long n) {
  for (long i = 0; i < n; ++i) { Vec3 ga; aa_to_rmat_bwd(Vec3{axis[3*i], axis[3*i+1], axis[3*i+2]}, ang[i], ld(G + 9 * i), &ga, gang + i); gaxis[3*i]=ga.x; gaxis[3*i+1]=ga.y; gaxis[3*i+2]=ga.z; }
}
void hm_expvec_bwd(const float* w, const float* G, float* gw, long n) {
  for (long i = 0; i < n; ++i) { Vec3 v{w[3*i], w[3*i+1], w[3*i+2]}; Vec3 g = exp_vec_bwd(v, exp_vec(v), ld(G + 9 * i)); gw[3*i]=g.x; gw[3*i+1]=g.y; gw[3*i+2]=g.z; }
}
}
extern "C" void hm_pair_kernel(const float* A, const float* B, float* out, long n, int gaussian) {
  for (long i = 0; i < n; ++i) {
    float qa[4], qb[4];
    rmat_to_quat(ld(A + 9 * i), qa);
    rmat_to_quat(ld(B + 9 * i), qb);
    const Quat a{qa[0], qa[1], qa[2], qa[3]}, b{qb[0], qb[1], qb[2], qb[3]};
    out[i] = gaussian ? so3_pair_kernel<true>(a, b) : so3_pair_kernel<false>(a, b);
  }
}
extern "C" void hm_normal4(unsigned long long seed, unsigned long long row0, unsigned long long offset, float* out, long n) {
  for (long i = 0; i < n; ++i) {
    const Normal4 z = normal4_from_u4(philox4x32_10(seed, row0 + i, offset));
    out[4*i] = z.a; out[4*i+1] = z.b; out[4*i+2] = z.c; out[4*i+3] = z.d;
  }
}
extern "C" void hm_logf_g(const float* w, const float* eps, float* logf, float* g, long n, int mode, int L) {
  for (long i = 0; i < n; ++i) igso3_logf_g(w[i], eps[i], mode, L, logf + i, g + i);
}

// ---- so3d_lanes.cuh: the one-lane and two-lane instantiations must equal the scalar functions they restate, bit for bit ----
#include "../../diffusion_extensions_b200/csrc/so3d_lanes.cuh"
static inline unsigned fbits(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
// returns the number of mismatching output words over n (pairs of) rotations; mean/x0h (n x 9) receive the L2 results
extern "C" long hm_lanes_p_mean(const float* x, const float* pred, const float* a, const float* b, const float* c1, const float* c2,
                                float* mean, float* x0h, long n, int has_pred) {
  long bad = 0;
  for (long i = 0; i + 1 < n; i += 2) {
    Mat3 r0 = ld(x + 9 * i), r1 = ld(x + 9 * (i + 1));
    // the step scalars are shared by the two lanes: use row i's
    Quat qh0, qh1;
    Quat q0, q1;
    if (has_pred) {
      q0 = p_mean_quat(r0, Vec3{pred[3*i], pred[3*i+1], pred[3*i+2]}, a[i], b[i], c1[i], c2[i], &qh0);
      q1 = p_mean_quat(r1, Vec3{pred[3*i+3], pred[3*i+4], pred[3*i+5]}, a[i], b[i], c1[i], c2[i], &qh1);
    } else {
      q0 = p_mean_quat_nopred(r0, a[i], c1[i], c2[i], &qh0);
      q1 = p_mean_quat_nopred(r1, a[i], c1[i], c2[i], &qh1);
    }
    const Mat3 m0 = quat_to_mat_unit(q0), m1 = quat_to_mat_unit(q1), h0 = quat_to_mat_unit(qh0), h1 = quat_to_mat_unit(qh1);
    // two lanes
    QuatL<L2> qh2;
    const Vec3L<L2> p2{L2{pred[3*i], pred[3*i+3]}, L2{pred[3*i+1], pred[3*i+4]}, L2{pred[3*i+2], pred[3*i+5]}};
    const QuatL<L2> q2 = has_pred ? p_mean_quat_l<L2, true>(lanes_of(r0, r1), p2, a[i], b[i], c1[i], c2[i], &qh2)
                                  : p_mean_quat_l<L2, false>(lanes_of(r0, r1), p2, a[i], b[i], c1[i], c2[i], &qh2);
    const Mat3L<L2> m2 = quat_to_mat_unit_l(q2), h2 = quat_to_mat_unit_l(qh2);
    // one lane
    QuatL<L1> qh1l;
    const Vec3L<L1> p1{L1{pred[3*i]}, L1{pred[3*i+1]}, L1{pred[3*i+2]}};
    const QuatL<L1> q1l = has_pred ? p_mean_quat_l<L1, true>(lanes_of(r0), p1, a[i], b[i], c1[i], c2[i], &qh1l)
                                   : p_mean_quat_l<L1, false>(lanes_of(r0), p1, a[i], b[i], c1[i], c2[i], &qh1l);
    const Mat3L<L1> m1l = quat_to_mat_unit_l(q1l);
    for (int k = 0; k < 9; ++k) {
      bad += fbits(m2.m[k].x) != fbits(m0.m[k]);
      bad += fbits(m2.m[k].y) != fbits(m1.m[k]);
      bad += fbits(h2.m[k].x) != fbits(h0.m[k]);
      bad += fbits(h2.m[k].y) != fbits(h1.m[k]);
      bad += fbits(m1l.m[k].x) != fbits(m0.m[k]);
      mean[9 * i + k] = m2.m[k].x; mean[9 * (i + 1) + k] = m2.m[k].y;
      x0h[9 * i + k] = h2.m[k].x; x0h[9 * (i + 1) + k] = h2.m[k].y;
    }
  }
  return bad;
}
// per-row schedule scalars (p_mean_quat_rows_l) against the scalar p_mean_quat: mismatching words
extern "C" long hm_lanes_p_mean_rows(const float* x, const float* pred, const float* a, const float* b, const float* c1, const float* c2, long n) {
  long bad = 0;
  for (long i = 0; i + 1 < n; i += 2) {
    const Mat3 r0 = ld(x + 9 * i), r1 = ld(x + 9 * (i + 1));
    Quat qh0, qh1;
    const Quat q0 = p_mean_quat(r0, Vec3{pred[3*i], pred[3*i+1], pred[3*i+2]}, a[i], b[i], c1[i], c2[i], &qh0);
    const Quat q1 = p_mean_quat(r1, Vec3{pred[3*i+3], pred[3*i+4], pred[3*i+5]}, a[i+1], b[i+1], c1[i+1], c2[i+1], &qh1);
    const Mat3 m0 = quat_to_mat_unit(q0), m1 = quat_to_mat_unit(q1);
    QuatL<L2> qh2;
    const Vec3L<L2> p2{L2{pred[3*i], pred[3*i+3]}, L2{pred[3*i+1], pred[3*i+4]}, L2{pred[3*i+2], pred[3*i+5]}};
    const QuatL<L2> q2 = p_mean_quat_rows_l<L2>(lanes_of(r0, r1), p2, L2{a[i], a[i+1]}, L2{b[i], b[i+1]}, L2{c1[i], c1[i+1]}, L2{c2[i], c2[i+1]}, &qh2);
    const Mat3L<L2> m2 = quat_to_mat_unit_l(q2);
    for (int k = 0; k < 9; ++k) {
      bad += fbits(m2.m[k].x) != fbits(m0.m[k]);
      bad += fbits(m2.m[k].y) != fbits(m1.m[k]);
    }
  }
  return bad;
}
// forward noising over lanes (q_sample_quat_l) against the one-row kernel's scalar sequence: mismatching words
extern "C" long hm_lanes_q_sample(const float* x, const float* sc, const float* axis, const float* ang, float* xt, long n) {
  long bad = 0;
  for (long i = 0; i + 1 < n; i += 2) {
    const Mat3 r0 = ld(x + 9 * i), r1 = ld(x + 9 * (i + 1));
    Mat3 m[2], nm[2];
    for (int j = 0; j < 2; ++j) {
      const long r = i + j;
      const Vec3 a{axis[3*r], axis[3*r+1], axis[3*r+2]};
      const Quat qn = quat_axis_angle(a, ang[r]);
      const AxisAngleF ax = axis_angle_fast(j ? r1 : r0);
      m[j] = quat_to_mat_unit(qmul(quat_axis_angle(ax.axis, sc[r] * ax.theta), qn));
      nm[j] = quat_to_mat_unit(qn);
    }
    const Vec3L<L2> a2{L2{axis[3*i], axis[3*i+3]}, L2{axis[3*i+1], axis[3*i+4]}, L2{axis[3*i+2], axis[3*i+5]}};
    QuatL<L2> qn2;
    const QuatL<L2> q2 = q_sample_quat_l<L2>(lanes_of(r0, r1), L2{sc[i], sc[i+1]}, a2, L2{ang[i], ang[i+1]}, &qn2);
    const Mat3L<L2> m2 = quat_to_mat_unit_l(q2), n2 = quat_to_mat_unit_l(qn2);
    const Vec3L<L1> a1{L1{axis[3*i]}, L1{axis[3*i+1]}, L1{axis[3*i+2]}};
    QuatL<L1> qn1;
    const Mat3L<L1> m1 = quat_to_mat_unit_l(q_sample_quat_l<L1>(lanes_of(r0), L1{sc[i]}, a1, L1{ang[i]}, &qn1));
    for (int k = 0; k < 9; ++k) {
      bad += fbits(m2.m[k].x) != fbits(m[0].m[k]);
      bad += fbits(m2.m[k].y) != fbits(m[1].m[k]);
      bad += fbits(n2.m[k].x) != fbits(nm[0].m[k]);
      bad += fbits(n2.m[k].y) != fbits(nm[1].m[k]);
      bad += fbits(m1.m[k].x) != fbits(m[0].m[k]);
      xt[9 * i + k] = m2.m[k].x; xt[9 * (i + 1) + k] = m2.m[k].y;
    }
  }
  return bad;
}
extern "C" long hm_lanes_sphere(const float* ua, const float* ub, float* axis, long n) {
  long bad = 0;
  for (long i = 0; i + 1 < n; i += 2) {
    const Vec3 a0 = sphere_from_uniforms(ua[i], ub[i]), a1 = sphere_from_uniforms(ua[i + 1], ub[i + 1]);
    const Vec3L<L2> v = sphere_from_uniforms_l(L2{ua[i], ua[i + 1]}, L2{ub[i], ub[i + 1]});
    bad += (fbits(v.x.x) != fbits(a0.x)) + (fbits(v.y.x) != fbits(a0.y)) + (fbits(v.z.x) != fbits(a0.z));
    bad += (fbits(v.x.y) != fbits(a1.x)) + (fbits(v.y.y) != fbits(a1.y)) + (fbits(v.z.y) != fbits(a1.z));
    axis[3*i] = v.x.x; axis[3*i+1] = v.y.x; axis[3*i+2] = v.z.x; axis[3*i+3] = v.x.y; axis[3*i+4] = v.y.y; axis[3*i+5] = v.z.y;
  }
  return bad;
}
// the warp-split series (one warp per rotation), emulated lane by lane with the kernel's reduction order
extern "C" void hm_series_warp(const float* w, const float* eps, float* logf, float* g, long n, int L, int guarded) {
  for (long i = 0; i < n; ++i) igso3_series_warp_host(w[i], eps[i], L, logf + i, g + i, guarded != 0);
}
