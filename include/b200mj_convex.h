/* b200mj_convex.h — narrow phase for the geom pairs that have no closed form: penetration of two convex primitives by
 * Minkowski Portal Refinement, and capsule-box through sphere-box tests along the capsule's segment.
 *
 * Stands where MuJoCo's `mjc_Convex` / `mjc_CapsuleBox` (called from `mj_collision`, reached by the reference through
 * mujoco.mj_step1 / mj_forward, dm_control/mujoco/engine.py:147-176) stand: ellipsoid / cylinder / box against
 * sphere / capsule / ellipsoid / cylinder / box — the quadruped's torso ellipsoid and eye cylinders against its legs
 * (dm_control/suite/quadruped.xml:95-97,122), the CMU walker's ellipsoid hands (locomotion/walkers/assets/
 * humanoid_CMU_V2019.xml:147,187) and capsules against the corridor's wall / platform boxes
 * (locomotion/arenas/corridors.py:394-440).
 *
 * MuJoCo <= 3.1 resolved these pairs with libccd's MPR (ccdMPRPenetration, tolerance 1e-6, 50 iterations) and newer
 * versions with their own GJK/EPA; neither library is in /root/reference, so this is a restatement of the published
 * MPR algorithm (G. Snethen, "XenoCollide", Game Programming Gems 7) with MuJoCo's conventions: one contact,
 * dist = -depth, normal from geom1 to geom2, position midway between the two witness points. PARITY UNPINNED for
 * these pairs: contact positions of real MuJoCo 3.11 may differ by its solver tolerance.
 *
 * Plain C++ with no dynamic indexing of local arrays: included by the CUDA engine (device code, one pair per lane) and
 * by the CPU oracle, so both sides run the same arithmetic; tests/test_convex_pairs.py checks it against the closed
 * forms it must reproduce (sphere-sphere, sphere-box, sphere-capsule, ...) and against geometric invariants.
 */
#ifndef B200MJ_CONVEX_H_
#define B200MJ_CONVEX_H_

#include <math.h>

#if defined(__CUDACC__) && !defined(B200MJ_CPU_EMU)
#define BMJ_HD __host__ __device__ __forceinline__
#define BMJ_HD_NOINLINE __host__ __device__ __noinline__
#else
#define BMJ_HD static inline
#define BMJ_HD_NOINLINE static
#endif
#if defined(__CUDACC__)
#define BMJ_UNROLL _Pragma("unroll")
#else
#define BMJ_UNROLL
#endif

#define BMJ_CCD_TOLERANCE 1e-6
#define BMJ_CCD_ITERATIONS 50

struct CvxGeom { int type; const double* pos; const double* mat; const double* size; double inflate; };
struct CvxSup { double v[3], a[3], b[3]; };      /* v = a - b: point of the Minkowski difference, a on geom1, b on geom2 */

BMJ_HD double cvx_dot(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
BMJ_HD void cvx_cross(double* r, const double* a, const double* b) {
  const double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
BMJ_HD double cvx_normalize(double* a) {
  const double n = sqrt(cvx_dot(a, a));
  if (n < 1e-300) { a[0] = 1; a[1] = 0; a[2] = 0; return 0; }
  const double inv = 1 / n;
  a[0] *= inv; a[1] *= inv; a[2] *= inv;
  return n;
}
BMJ_HD double cvx_sign(double x) { return x < 0 ? -1.0 : 1.0; }

/* support point of one primitive (world frame) in the UNIT direction d; geom types: b200mj_model_fields.h */
BMJ_HD void cvx_support1(const CvxGeom& g, const double* d, double* out) {
  const double* m = g.mat; const double* s = g.size;
  /* direction in the geom frame */
  const double lx = m[0]*d[0] + m[3]*d[1] + m[6]*d[2], ly = m[1]*d[0] + m[4]*d[1] + m[7]*d[2], lz = m[2]*d[0] + m[5]*d[1] + m[8]*d[2];
  double px = 0, py = 0, pz = 0, ball = g.inflate;
  if (g.type == 2) ball += s[0];                                   /* sphere */
  else if (g.type == 3) { pz = cvx_sign(lz) * s[1]; ball += s[0]; } /* capsule: segment + ball */
  else if (g.type == 4) {                                           /* ellipsoid */
    const double ax = s[0]*s[0]*lx, ay = s[1]*s[1]*ly, az = s[2]*s[2]*lz;
    const double den = sqrt(ax*lx + ay*ly + az*lz);
    if (den > 1e-300) { px = ax / den; py = ay / den; pz = az / den; }
  } else if (g.type == 5) {                                         /* cylinder */
    const double rr = sqrt(lx*lx + ly*ly);
    if (rr > 1e-300) { px = s[0] * lx / rr; py = s[0] * ly / rr; }
    pz = cvx_sign(lz) * s[1];
  } else {                                                          /* box */
    px = cvx_sign(lx) * s[0]; py = cvx_sign(ly) * s[1]; pz = cvx_sign(lz) * s[2];
  }
  out[0] = g.pos[0] + m[0]*px + m[1]*py + m[2]*pz + ball * d[0];
  out[1] = g.pos[1] + m[3]*px + m[4]*py + m[5]*pz + ball * d[1];
  out[2] = g.pos[2] + m[6]*px + m[7]*py + m[8]*pz + ball * d[2];
}

BMJ_HD void cvx_support(const CvxGeom& g1, const CvxGeom& g2, const double* dir, CvxSup& s) {
  const double nd[3] = {-dir[0], -dir[1], -dir[2]};
  cvx_support1(g1, dir, s.a);
  cvx_support1(g2, nd, s.b);
  s.v[0] = s.a[0] - s.b[0]; s.v[1] = s.a[1] - s.b[1]; s.v[2] = s.a[2] - s.b[2];
}

/* squared distance from the origin to triangle (a, b, c); witness = closest point */
BMJ_HD double cvx_origin_tri_dist2(const double* a, const double* b, const double* c, double* witness) {
  double d1[3] = {b[0]-a[0], b[1]-a[1], b[2]-a[2]}, d2[3] = {c[0]-a[0], c[1]-a[1], c[2]-a[2]};
  const double u = cvx_dot(a, a), v = cvx_dot(d1, d1), w = cvx_dot(d2, d2);
  const double p = cvx_dot(a, d1), q = cvx_dot(a, d2), r = cvx_dot(d1, d2);
  const double den = w * v - r * r;
  double s = -1, t = -1;
  if (fabs(den) > 1e-300) { s = (q * r - w * p) / den; t = (-s * r - q) / w; }
  if (s > -1e-15 && s < 1 + 1e-15 && t > -1e-15 && t < 1 + 1e-15 && t + s < 1 + 1e-15 && fabs(den) > 1e-300) {
    for (int i = 0; i < 3; i++) witness[i] = a[i] + s * d1[i] + t * d2[i];
    double dist = s*s*v + t*t*w + 2*s*t*r + 2*s*p + 2*t*q + u;
    return dist < 0 ? 0 : dist;
  }
  /* closest point lies on an edge: test the three segments */
  double best = 1e300;
  BMJ_UNROLL
  for (int e = 0; e < 3; e++) {
    const double* x0 = e == 0 ? a : (e == 1 ? a : b);
    const double* x1 = e == 0 ? b : (e == 1 ? c : c);
    const double dd[3] = {x1[0]-x0[0], x1[1]-x0[1], x1[2]-x0[2]};
    const double len2 = cvx_dot(dd, dd);
    double tt = len2 > 1e-300 ? -cvx_dot(x0, dd) / len2 : 0.0;
    tt = tt < 0 ? 0 : (tt > 1 ? 1 : tt);
    const double pt[3] = {x0[0] + tt*dd[0], x0[1] + tt*dd[1], x0[2] + tt*dd[2]};
    const double dist = cvx_dot(pt, pt);
    if (dist < best) { best = dist; witness[0] = pt[0]; witness[1] = pt[1]; witness[2] = pt[2]; }
  }
  return best;
}

/* Penetration of two convex primitives. Returns 1 and (depth >= 0, unit dir from geom1 to geom2, pos) when they
 * overlap, 0 otherwise. */
BMJ_HD_NOINLINE int cvx_mpr(const CvxGeom& g1, const CvxGeom& g2, const double* interior, double* depth, double* dir_out,
                            double* pos_out) {
  CvxSup v0, v1, v2, v3, v4;
  double dir[3], va[3], vb[3];
  /* interior point of the Minkowski difference: the difference of the centres, or the caller's (refinement passes) */
  for (int i = 0; i < 3; i++) {
    const double mid = 0.5 * (g1.pos[i] + g2.pos[i]);
    v0.v[i] = interior ? interior[i] : g1.pos[i] - g2.pos[i];
    v0.a[i] = mid + 0.5 * v0.v[i]; v0.b[i] = mid - 0.5 * v0.v[i];
  }
  if (cvx_dot(v0.v, v0.v) < 1e-20) v0.v[0] += 1e-5;
  /* ---- discover a portal ---- */
  dir[0] = -v0.v[0]; dir[1] = -v0.v[1]; dir[2] = -v0.v[2]; cvx_normalize(dir);
  cvx_support(g1, g2, dir, v1);
  if (cvx_dot(v1.v, dir) <= 0) return 0;
  cvx_cross(dir, v0.v, v1.v);
  if (cvx_dot(dir, dir) < 1e-20) {
    /* origin on the ray v0 -> v1: the penetration is along it */
    double d[3] = {v1.v[0], v1.v[1], v1.v[2]};
    *depth = cvx_normalize(d);
    for (int i = 0; i < 3; i++) { dir_out[i] = d[i]; pos_out[i] = 0.5 * (v1.a[i] + v1.b[i]); }
    return 1;
  }
  cvx_normalize(dir);
  cvx_support(g1, g2, dir, v2);
  if (cvx_dot(v2.v, dir) <= 0) return 0;
  for (int i = 0; i < 3; i++) { va[i] = v1.v[i] - v0.v[i]; vb[i] = v2.v[i] - v0.v[i]; }
  cvx_cross(dir, va, vb); cvx_normalize(dir);
  if (cvx_dot(dir, v0.v) > 0) { CvxSup t = v1; v1 = v2; v2 = t; dir[0] = -dir[0]; dir[1] = -dir[1]; dir[2] = -dir[2]; }
  for (int it = 0; ; it++) {
    if (it > BMJ_CCD_ITERATIONS) return 0;
    cvx_support(g1, g2, dir, v3);
    if (cvx_dot(v3.v, dir) <= 0) return 0;
    int cont = 0;
    cvx_cross(va, v1.v, v3.v);
    if (cvx_dot(va, v0.v) < 0) { v2 = v3; cont = 1; }
    if (!cont) { cvx_cross(va, v3.v, v2.v); if (cvx_dot(va, v0.v) < 0) { v1 = v3; cont = 1; } }
    if (!cont) break;
    for (int i = 0; i < 3; i++) { va[i] = v1.v[i] - v0.v[i]; vb[i] = v2.v[i] - v0.v[i]; }
    cvx_cross(dir, va, vb); cvx_normalize(dir);
  }
  /* ---- does the origin lie inside? move the portal outwards until it passes the origin (overlap) or the support
   * plane shows that it never will (separated) ---- */
  for (int it = 0; ; it++) {
    for (int i = 0; i < 3; i++) { va[i] = v2.v[i] - v1.v[i]; vb[i] = v3.v[i] - v1.v[i]; }
    cvx_cross(dir, va, vb); cvx_normalize(dir);
    if (cvx_dot(dir, v1.v) >= 0) break;                     /* the portal encloses the origin */
    cvx_support(g1, g2, dir, v4);
    const double d4 = cvx_dot(v4.v, dir);
    if (d4 < 0) return 0;                                   /* origin beyond the support plane: separated */
    double reach = d4 - cvx_dot(v1.v, dir);
    const double r2 = d4 - cvx_dot(v2.v, dir), r3 = d4 - cvx_dot(v3.v, dir);
    if (r2 < reach) reach = r2;
    if (r3 < reach) reach = r3;
    if (reach <= BMJ_CCD_TOLERANCE || it >= BMJ_CCD_ITERATIONS) return 0;
    double v4v0[3]; cvx_cross(v4v0, v4.v, v0.v);
    if (cvx_dot(v1.v, v4v0) > 0) { if (cvx_dot(v2.v, v4v0) > 0) v1 = v4; else v3 = v4; }
    else { if (cvx_dot(v3.v, v4v0) > 0) v2 = v4; else v1 = v4; }
  }
  /* ---- refine it towards the surface of the Minkowski difference ---- */
  for (int it = 0; ; it++) {
    for (int i = 0; i < 3; i++) { va[i] = v2.v[i] - v1.v[i]; vb[i] = v3.v[i] - v1.v[i]; }
    cvx_cross(dir, va, vb); cvx_normalize(dir);
    cvx_support(g1, g2, dir, v4);
    const double d4 = cvx_dot(v4.v, dir);
    double reach = d4 - cvx_dot(v1.v, dir);
    const double r2 = d4 - cvx_dot(v2.v, dir), r3 = d4 - cvx_dot(v3.v, dir);
    if (r2 < reach) reach = r2;
    if (r3 < reach) reach = r3;
    if (reach <= BMJ_CCD_TOLERANCE || it >= BMJ_CCD_ITERATIONS) break;
    /* expand the portal with v4: which of v1..v3 does it replace? */
    double v4v0[3]; cvx_cross(v4v0, v4.v, v0.v);
    if (cvx_dot(v1.v, v4v0) > 0) { if (cvx_dot(v2.v, v4v0) > 0) v1 = v4; else v3 = v4; }
    else { if (cvx_dot(v3.v, v4v0) > 0) v2 = v4; else v1 = v4; }
  }
  double wit[3] = {0, 0, 0};
  *depth = sqrt(cvx_origin_tri_dist2(v1.v, v2.v, v3.v, wit));
  if (cvx_dot(wit, wit) < 1e-30) { wit[0] = dir[0]; wit[1] = dir[1]; wit[2] = dir[2]; }
  cvx_normalize(wit);
  for (int i = 0; i < 3; i++) dir_out[i] = wit[i];
  /* position: barycentric combination of the witness points (libccd's findPos) */
  double b0, b1, b2, b3, tmp[3];
  cvx_cross(tmp, v1.v, v2.v); b0 = cvx_dot(tmp, v3.v);
  cvx_cross(tmp, v3.v, v2.v); b1 = cvx_dot(tmp, v0.v);
  cvx_cross(tmp, v0.v, v1.v); b2 = cvx_dot(tmp, v3.v);
  cvx_cross(tmp, v2.v, v1.v); b3 = cvx_dot(tmp, v0.v);
  double sum = b0 + b1 + b2 + b3;
  if (sum <= 0) {
    b0 = 0;
    cvx_cross(tmp, v2.v, v3.v); b1 = cvx_dot(tmp, dir);
    cvx_cross(tmp, v3.v, v1.v); b2 = cvx_dot(tmp, dir);
    cvx_cross(tmp, v1.v, v2.v); b3 = cvx_dot(tmp, dir);
    sum = b1 + b2 + b3;
  }
  const double inv = fabs(sum) > 1e-300 ? 1 / sum : 0.0;
  for (int i = 0; i < 3; i++) {
    const double pa = (b0 * v0.a[i] + b1 * v1.a[i] + b2 * v2.a[i] + b3 * v3.a[i]) * inv;
    const double pb = (b0 * v0.b[i] + b1 * v1.b[i] + b2 * v2.b[i] + b3 * v3.b[i]) * inv;
    pos_out[i] = 0.5 * (pa + pb);
  }
  return 1;
}

/* One contact of a convex pair (MuJoCo's mjc_Convex conventions). margin > 0 inflates both shapes by margin / 2.
 * Returns the number of contacts (0 or 1); out = dist, pos[3], normal[3] (geom1 -> geom2). */
BMJ_HD int cvx_pair(int t1, const double* p1, const double* m1, const double* s1, int t2, const double* p2, const double* m2,
                    const double* s2, double margin, double* dist, double* pos, double* nrm) {
  CvxGeom g1 = {t1, p1, m1, s1, 0.5 * margin}, g2 = {t2, p2, m2, s2, 0.5 * margin};
  double depth;
  if (!cvx_mpr(g1, g2, (const double*)0, &depth, nrm, pos)) return 0;
#ifdef BMJ_CVX_SKIP_REFINE      /* timing experiments only: what the refinement below costs (results differ) */
  *dist = margin - depth;
  return 1;
#endif
  /* MPR measures the overlap along the ray from its interior point through the origin, which is the minimum
   * translation (what GJK/EPA returns) only when that ray is parallel to the contact normal. So: (1) walk the direction
   * downhill on h(d) = max over the Minkowski difference of x.d — the overlap along d, whose minimum over unit d is the
   * penetration depth and whose tangential gradient is the support point — with a step sized from the local extent of
   * the difference and halved until h decreases; (2) run MPR once more from an interior point placed behind the origin
   * on that direction, which yields depth, normal and position consistently. */
  /* h has one local minimum per face / edge region of a box-like difference and MPR lands in the one its centre ray
   * points at. Three walks, from MPR's direction, from the best frame axis of the two geoms and from the best point of
   * a coarse Fibonacci lattice of the sphere; the lowest end point wins. */
  double start[9] = {nrm[0], nrm[1], nrm[2], nrm[0], nrm[1], nrm[2], nrm[0], nrm[1], nrm[2]};
  {
    double best = 1e300;
    for (int k = 0; k < 12; k++) {
      const double* mm = k < 6 ? m1 : m2;
      const int ax = (k % 6) >> 1; const double sg = (k & 1) ? -1.0 : 1.0;
      const double dk[3] = {sg * mm[ax], sg * mm[3 + ax], sg * mm[6 + ax]};
      CvxSup sk;
      cvx_support(g1, g2, dk, sk);
      const double hk = cvx_dot(sk.v, dk);
      if (hk < best) { best = hk; start[3] = dk[0]; start[4] = dk[1]; start[5] = dk[2]; }
    }
    best = 1e300;
    for (int k = 0; k < 96; k++) {
      const double z = 1.0 - (2.0 * k + 1.0) / 96.0, rr = sqrt(1.0 - z * z), ph = 2.399963229728653 * k;
      const double dk[3] = {rr * cos(ph), rr * sin(ph), z};
      CvxSup sk;
      cvx_support(g1, g2, dk, sk);
      const double hk = cvx_dot(sk.v, dk);
      if (hk < best) { best = hk; start[6] = dk[0]; start[7] = dk[1]; start[8] = dk[2]; }
    }
  }
  double dbest[3] = {nrm[0], nrm[1], nrm[2]}, hbest = 1e300;
  CvxSup sbest;
  cvx_support(g1, g2, dbest, sbest);
  for (int c = 0; c < 3; c++) {
    double d[3] = {start[3 * c], start[3 * c + 1], start[3 * c + 2]};
    CvxSup sf;
    cvx_support(g1, g2, d, sf);
    double hd = cvx_dot(sf.v, d), eta = 0;
    for (int it = 0; it < 32; it++) {
      const double g[3] = {sf.v[0] - hd * d[0], sf.v[1] - hd * d[1], sf.v[2] - hd * d[2]};
      if (cvx_dot(g, g) < 1e-20) break;
      if (it == 0) {      /* first step: sized from the extent of the difference along d (exact for a ball) */
        const double nd[3] = {-d[0], -d[1], -d[2]};
        CvxSup sb;
        cvx_support(g1, g2, nd, sb);
        const double hb = cvx_dot(sb.v, nd);
        eta = 2.0 / (hb - hd > 1e-9 ? hb - hd : 1e-9);
      }
      int accepted = 0;
      for (int half = 0; half < 8 && !accepted; half++) {
        double dn[3] = {d[0] - eta * g[0], d[1] - eta * g[1], d[2] - eta * g[2]};
        cvx_normalize(dn);
        CvxSup sn;
        cvx_support(g1, g2, dn, sn);
        const double hn = cvx_dot(sn.v, dn);
        if (hn < hd - 1e-14) {
          /* Barzilai-Borwein step for the next iteration from the change of direction and of tangential gradient */
          const double gn[3] = {sn.v[0] - hn * dn[0], sn.v[1] - hn * dn[1], sn.v[2] - hn * dn[2]};
          const double dd[3] = {dn[0] - d[0], dn[1] - d[1], dn[2] - d[2]}, dg[3] = {gn[0] - g[0], gn[1] - g[1], gn[2] - g[2]};
          const double sy = cvx_dot(dd, dg), ss = cvx_dot(dd, dd);
          d[0] = dn[0]; d[1] = dn[1]; d[2] = dn[2]; sf = sn; hd = hn; accepted = 1;
          eta = (sy > 1e-12 * ss && ss > 0) ? ss / sy : 2 * eta;
        } else eta *= 0.5;
      }
      if (!accepted) break;
    }
    if (hd < hbest) { hbest = hd; dbest[0] = d[0]; dbest[1] = d[1]; dbest[2] = d[2]; sbest = sf; }
  }
  if (hbest < depth - 1e-9) {
    /* MPR once more from behind the origin on the winning direction: on a flat face of the difference (boxes, cylinder
     * caps) its portal lies in the face and returns the face normal, which the (sub)gradient walk cannot reach at a
     * kink of h; otherwise the walk's own answer stands (support plane along the direction) */
    const double nd[3] = {-dbest[0], -dbest[1], -dbest[2]};
    CvxSup sb;
    cvx_support(g1, g2, nd, sb);
    const double ext = cvx_dot(sb.v, nd);
    double d2 = 1e300, n2[3], p2v[3];
    int hit = 0;
    if (ext > 1e-12) {
      const double inner[3] = {0.5 * ext * nd[0], 0.5 * ext * nd[1], 0.5 * ext * nd[2]};
      hit = cvx_mpr(g1, g2, inner, &d2, n2, p2v);
    }
    if (hit && d2 <= hbest + 1e-9) {
      depth = d2;
      for (int i = 0; i < 3; i++) { nrm[i] = n2[i]; pos[i] = p2v[i]; }
    } else {
      depth = hbest;
      for (int i = 0; i < 3; i++) { nrm[i] = dbest[i]; pos[i] = 0.5 * (sbest.a[i] + sbest.b[i]); }
    }
  }
  *dist = margin - depth;
  return 1;
}

/* sphere (centre c, radius r) against a box given in ITS OWN frame (half sizes s): signed distance, local normal
 * (sphere -> box), closest-feature logic identical to the sphere-box pair of the engine. Returns 0 when farther than margin. */
BMJ_HD int cvx_sphere_box_local(const double* c, double r, const double* s, double margin, double* dist, double* nl) {
  double cl[3]; int inside = 1;
  for (int i = 0; i < 3; i++) { cl[i] = c[i] < -s[i] ? -s[i] : (c[i] > s[i] ? s[i] : c[i]); if (cl[i] != c[i]) inside = 0; }
  if (!inside) {
    const double dl[3] = {c[0] - cl[0], c[1] - cl[1], c[2] - cl[2]};
    const double dn = sqrt(cvx_dot(dl, dl));
    if (dn - r > margin) return 0;
    *dist = dn - r;
    for (int i = 0; i < 3; i++) nl[i] = -dl[i] / dn;
  } else {
    double bd = s[0] - fabs(c[0]); int best = 0;
    if (s[1] - fabs(c[1]) < bd) { bd = s[1] - fabs(c[1]); best = 1; }
    if (s[2] - fabs(c[2]) < bd) { bd = s[2] - fabs(c[2]); best = 2; }
    nl[0] = best == 0 ? (c[0] > 0 ? -1.0 : 1.0) : 0.0;
    nl[1] = best == 1 ? (c[1] > 0 ? -1.0 : 1.0) : 0.0;
    nl[2] = best == 2 ? (c[2] > 0 ? -1.0 : 1.0) : 0.0;
    *dist = -bd - r;
  }
  return 1;
}

/* squared distance from point q (box frame) to the box */
BMJ_HD double cvx_point_box_dist2(const double* q, const double* s) {
  double d2 = 0;
  for (int i = 0; i < 3; i++) { const double e = fabs(q[i]) - s[i]; if (e > 0) d2 += e * e; }
  return d2;
}

/* Capsule (geom1: centre p1, axis = column 2 of m1, radius s1[0], half length s1[1]) against a box (geom2).
 * Up to two contacts, each a sphere-box test at a point of the capsule's segment: the point of the segment closest to
 * the box (found by golden-section search on the convex distance function; skipped when the distance is flat along
 * the segment, i.e. the capsule lies parallel to a face) and the end points. out: [k*7] = dist, pos[3], normal[3]
 * (capsule -> box). Returns the number of contacts. */
BMJ_HD_NOINLINE int cvx_capsule_box(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                                    const double* s2, double margin, double* out) {
  const double r = s1[0], hl = s1[1];
  /* segment in the box frame: c + t * ax, t in [-hl, hl] */
  const double w[3] = {p1[0] - p2[0], p1[1] - p2[1], p1[2] - p2[2]}, axw[3] = {m1[2], m1[5], m1[8]};
  double c[3], ax[3];
  for (int i = 0; i < 3; i++) { c[i] = m2[i]*w[0] + m2[3+i]*w[1] + m2[6+i]*w[2]; ax[i] = m2[i]*axw[0] + m2[3+i]*axw[1] + m2[6+i]*axw[2]; }
  /* golden-section minimisation of the (convex) distance to the box along the segment */
  double lo = -hl, hi = hl;
  const double gr = 0.6180339887498949;
  double x1 = hi - gr * (hi - lo), x2 = lo + gr * (hi - lo);
  double q[3];
  for (int i = 0; i < 3; i++) q[i] = c[i] + x1 * ax[i];
  double f1 = cvx_point_box_dist2(q, s2);
  for (int i = 0; i < 3; i++) q[i] = c[i] + x2 * ax[i];
  double f2 = cvx_point_box_dist2(q, s2);
  for (int it = 0; it < 48; it++) {
    if (f1 <= f2) { hi = x2; x2 = x1; f2 = f1; x1 = hi - gr * (hi - lo); for (int i = 0; i < 3; i++) q[i] = c[i] + x1 * ax[i]; f1 = cvx_point_box_dist2(q, s2); }
    else { lo = x1; x1 = x2; f1 = f2; x2 = lo + gr * (hi - lo); for (int i = 0; i < 3; i++) q[i] = c[i] + x2 * ax[i]; f2 = cvx_point_box_dist2(q, s2); }
  }
  const double tmid = 0.5 * (lo + hi);
  /* distances at the two ends and at the minimiser */
  double qa[3], qb[3], qm[3];
  for (int i = 0; i < 3; i++) { qa[i] = c[i] - hl * ax[i]; qb[i] = c[i] + hl * ax[i]; qm[i] = c[i] + tmid * ax[i]; }
  const double da = cvx_point_box_dist2(qa, s2), db = cvx_point_box_dist2(qb, s2), dm = cvx_point_box_dist2(qm, s2);
  /* an interior minimiser counts only when it is strictly better than both ends (otherwise the ends describe the contact) */
  const double eps = 1e-12 * (1 + da + db);
  const int use_mid = (tmid > -hl + 1e-9 && tmid < hl - 1e-9 && dm < da - eps && dm < db - eps) ? 1 : 0;
  int n = 0;
  double dist, nl[3];
  /* candidate order: end -hl, interior, end +hl; at most two are kept: the interior point displaces the shallower end */
  double cd[3] = {1e300, 1e300, 1e300}, cn[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}; int ok[3] = {0, 0, 0};
  if (cvx_sphere_box_local(qa, r, s2, margin, &dist, nl)) { ok[0] = 1; cd[0] = dist; cn[0] = nl[0]; cn[1] = nl[1]; cn[2] = nl[2]; }
  if (use_mid && cvx_sphere_box_local(qm, r, s2, margin, &dist, nl)) { ok[1] = 1; cd[1] = dist; cn[3] = nl[0]; cn[4] = nl[1]; cn[5] = nl[2]; }
  if (cvx_sphere_box_local(qb, r, s2, margin, &dist, nl)) { ok[2] = 1; cd[2] = dist; cn[6] = nl[0]; cn[7] = nl[1]; cn[8] = nl[2]; }
  if (ok[0] && ok[1] && ok[2]) { if (cd[0] <= cd[2]) ok[2] = 0; else ok[0] = 0; }
  BMJ_UNROLL
  for (int k = 0; k < 3; k++) {
    if (!ok[k]) continue;
    const double* qq = k == 0 ? qa : (k == 1 ? qm : qb);
    const double* nk = cn + 3 * k;
    double* o = out + 7 * n;
    o[0] = cd[k];
    for (int i = 0; i < 3; i++) {
      const double nw = m2[3*i] * nk[0] + m2[3*i+1] * nk[1] + m2[3*i+2] * nk[2];          /* normal, world frame */
      const double cw = p2[i] + m2[3*i] * qq[0] + m2[3*i+1] * qq[1] + m2[3*i+2] * qq[2];  /* sphere centre, world frame */
      o[4 + i] = nw;
      o[1 + i] = cw + nw * (r + 0.5 * cd[k]);
    }
    n++;
  }
  return n;
}

#endif /* B200MJ_CONVEX_H_ */
