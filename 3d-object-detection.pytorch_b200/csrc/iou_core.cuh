// 2D-keypoint based 3D IoU, one sample pair per thread, double precision.
// Replaces the per-sample CPU loop of the reference's last-epoch evaluation:
//   torchdet3d/utils/geometry.py:51-108        lift_2d      (EPnP-style lift: 16 x 12 system, smallest eigenvector)
//   torchdet3d/evaluation/metrics.py:70-89     compute_2d_based_iou
//   3rdparty/Objectron/objectron/dataset/box.py:123-156, 207-225   Box.fit, Box.volume
//   3rdparty/Objectron/objectron/dataset/iou.py:22-35, 74-211       IoU.iou (Sutherland-Hodgman clipping)
// The work is tiny and branchy (a 12 x 12 symmetric eigen-solve and 12 quad-against-box clippings per pair), so the design
// goal is only "no host round trip and no Python loop": every function here is plain C++ (`TD3D_HD`), which lets
// tests/host/iou_emul.cpp run the very same code on the CPU against the oracle.
//
// Differences from the reference's numerics, all below the test tolerance: the eigen-solve is cyclic Jacobi instead of
// LAPACK's tridiagonal QR; the least-squares box fit uses the closed form of its diagonal normal equations; the
// hull volume of the intersection points is the sum of the pyramids the clipped faces close over an interior point (the
// reference calls Qhull on the same points); a brute-force hull takes over when faces of the two boxes (nearly) coincide.
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define TD3D_HD __host__ __device__ __forceinline__
#else
#define TD3D_HD inline
#endif

namespace td3d {
namespace iou3d {

// EPnP barycentric weights of the 8 box vertices w.r.t. the 4 control points (geometry.py:6-13)
TD3D_HD double epnp_alpha(int i, int j) {
  const int a[8][4] = {{4, -1, -1, -1}, {2, -1, -1, 1}, {2, -1, 1, -1}, {0, -1, 1, 1}, {2, 1, -1, -1}, {0, 1, -1, 1}, {0, 1, 1, -1}, {-2, 1, 1, 1}};
  return (double)a[i][j];
}

// Smallest-eigenvalue eigenvector of a symmetric 12 x 12 matrix (cyclic Jacobi, eigenvectors accumulated in v).
TD3D_HD void smallest_eigvec12(double a[12][12], double vec[12]) {
  double v[12][12];
  for (int i = 0; i < 12; ++i)
    for (int j = 0; j < 12; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  // The wanted eigenvalue is the (near-zero) smallest one and its gap to the next is small against the norm of the matrix,
  // so the sweeps run until the off-diagonal mass stops shrinking, not until it is small against the diagonal.
  // The three loops must stay rolled: fully unrolled (66 rotations with constant p, q) nvcc 12.9 -O3 produced device code
  // whose rotations were not similarity transforms (trace 461 -> 229 on the known-answer keypoints), while the host build
  // of the same source was exact.
  double prev_off = -1.0;
#pragma unroll 1
  for (int sweep = 0; sweep < 60; ++sweep) {
    double off = 0.0;
    for (int i = 0; i < 12; ++i)
      for (int j = i + 1; j < 12; ++j) off += a[i][j] * a[i][j];
    if (off == 0.0 || (prev_off >= 0.0 && off >= prev_off)) break;
    prev_off = off;
#pragma unroll 1
    for (int p = 0; p < 11; ++p)
#pragma unroll 1
      for (int q = p + 1; q < 12; ++q) {
        const double apq = a[p][q];
        if (apq == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
        const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 12; ++k) {          // A <- A J  (columns p, q)
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 12; ++k) {          // A <- J^T A  (rows p, q)
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 12; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 12; ++i)
    if (a[i][i] < a[best][best]) best = i;
  for (int k = 0; k < 12; ++k) vec[k] = v[k][best];
}

// geometry.py:64-106 for one set of nine (x, y) keypoints in [0, 1]; cam = {fx, fy, cx, cy} of the NDC camera matrix
TD3D_HD void lift_2d(const float* kp, int portrait, const double cam[4], double out[9][3]) {
  double m[16][12];
  for (int r = 0; r < 16; ++r)
    for (int c = 0; c < 12; ++c) m[r][c] = 0.0;
  for (int i = 0; i < 8; ++i) {
    // u, v in float32: the reference forms them in the dtype of its input, which is float32 for model outputs (metrics.py:75-76)
    const float px = kp[2 * (i + 1)], py = kp[2 * (i + 1) + 1];
    const float uf = portrait ? py * 2.f - 1.f : px * 2.f - 1.f;
    const float wf = portrait ? px * 2.f - 1.f : 1.f - py * 2.f;
    const double u = (double)uf, w = (double)wf;
    for (int j = 0; j < 4; ++j) {
      const double al = epnp_alpha(i, j);
      m[2 * i][3 * j] = cam[0] * al;
      m[2 * i][3 * j + 2] = (cam[2] + u) * al;
      m[2 * i + 1][3 * j + 1] = cam[1] * al;
      m[2 * i + 1][3 * j + 2] = (cam[3] + w) * al;
    }
  }
  double a[12][12];
  for (int i = 0; i < 12; ++i)
    for (int j = i; j < 12; ++j) {
      double s = 0.0;
      for (int r = 0; r < 16; ++r) s += m[r][i] * m[r][j];
      a[i][j] = s;
      a[j][i] = s;
    }
  double ev[12];
  smallest_eigvec12(a, ev);
  const double sign = ev[2] > 0.0 ? -1.0 : 1.0;        // all points in front of the camera: control point 0 has z <= 0
  for (int c = 0; c < 3; ++c) out[0][c] = sign * ev[c];
  for (int i = 0; i < 8; ++i)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int j = 0; j < 4; ++j) s += epnp_alpha(i, j) * ev[3 * j + c];
      out[i + 1][c] = sign * s;
    }
}

struct Box {
  double r[3][3];      // fitted linear map (columns = box axes; not necessarily orthonormal)
  double t[3], s[3];   // translation, scale (mean edge length per axis)
  double vol;          // |det| of the three edges at vertex 1 of the ORIGINAL vertices
};

TD3D_HD void unit_corner(int i, double c[3]) {      // vertex i of the unit box (box.py:24-34): 0 = centre
  if (i == 0) { c[0] = c[1] = c[2] = 0.0; return; }
  const int k = i - 1;
  c[0] = (k & 4) ? 0.5 : -0.5;
  c[1] = (k & 2) ? 0.5 : -0.5;
  c[2] = (k & 1) ? 0.5 : -0.5;
}

// box.py:123-156 (fit) and :207-225 (volume)
TD3D_HD void box_fit(const double v[9][3], Box& b) {
  const int edges[12][2] = {{1, 5}, {2, 6}, {3, 7}, {4, 8}, {1, 3}, {5, 7}, {2, 4}, {6, 8}, {1, 2}, {3, 4}, {5, 6}, {7, 8}};
  for (int ax = 0; ax < 3; ++ax) {
    double s = 0.0;
    for (int e = 0; e < 4; ++e) {
      const int p = edges[ax * 4 + e][0], q = edges[ax * 4 + e][1];
      const double dx = v[p][0] - v[q][0], dy = v[p][1] - v[q][1], dz = v[p][2] - v[q][2];
      s += sqrt(dx * dx + dy * dy + dz * dz);
    }
    b.s[ax] = s / 4.0;
  }
  // least squares of [x | 1] sol = v with x the scaled axis-aligned vertices: the normal equations are diagonal
  // (sum x_a = 0, sum x_a x_b = 0, sum x_a^2 = 2 s_a^2, 9 ones), so sol[a] = sum_i x_ia v_i / (2 s_a^2), sol[3] = mean v
  for (int c = 0; c < 3; ++c) {
    double mean = 0.0;
    for (int i = 0; i < 9; ++i) mean += v[i][c];
    b.t[c] = mean / 9.0;
  }
  for (int ax = 0; ax < 3; ++ax)
    for (int c = 0; c < 3; ++c) {
      double s = 0.0;
      for (int i = 1; i < 9; ++i) {
        double u[3];
        unit_corner(i, u);
        s += u[ax] * b.s[ax] * v[i][c];
      }
      const double den = 2.0 * b.s[ax] * b.s[ax];
      b.r[c][ax] = den > 0.0 ? s / den : 0.0;          // orientation = sol[:3, :3]^T
    }
  const double i0 = v[2][0] - v[1][0], i1 = v[2][1] - v[1][1], i2 = v[2][2] - v[1][2];
  const double j0 = v[3][0] - v[1][0], j1 = v[3][1] - v[1][1], j2 = v[3][2] - v[1][2];
  const double k0 = v[5][0] - v[1][0], k1 = v[5][1] - v[1][1], k2 = v[5][2] - v[1][2];
  b.vol = fabs(i0 * (j1 * k2 - j2 * k1) - i1 * (j0 * k2 - j2 * k0) + i2 * (j0 * k1 - j1 * k0));
}

TD3D_HD bool inv3(const double m[3][3], double o[3][3]) {
  const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c01 = m[1][2] * m[2][0] - m[1][0] * m[2][2],
               c02 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  const double det = m[0][0] * c00 + m[0][1] * c01 + m[0][2] * c02;
  if (det == 0.0 || !(fabs(det) > 0.0)) return false;
  const double id = 1.0 / det;
  o[0][0] = c00 * id; o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][0] = c01 * id; o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][0] = c02 * id; o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return true;
}

struct Poly { double p[16][3]; int n; };

TD3D_HD int classify(double x, double plane, double normal) {      // iou.py:190-207 (thick plane)
  const double d = normal * (x - plane);
  return d > 1e-6 ? 1 : (d < -1e-6 ? -1 : 0);
}

// iou.py:99-157: Sutherland-Hodgman against the plane x[axis] = plane (points with normal * (x - plane) > 0 are kept).
// Returns true when every vertex lies ON the plane (the polygon is then left as it is).
TD3D_HD bool clip_poly(Poly& poly, double plane, double normal, int axis, bool& touched) {
  if (poly.n <= 1) { poly.n = 0; return false; }
  Poly res;
  res.n = 0;
  bool in_plane = true;
  for (int i = 0; i < poly.n; ++i) {
    const double* cur = poly.p[i];
    const double* prev = poly.p[(i + poly.n - 1) % poly.n];
    const int d1 = classify(prev[axis], plane, normal), d2 = classify(cur[axis], plane, normal);
    if (d2 == 0) touched = true;             // a vertex inside the thick plane: the configuration is (near-)degenerate
    bool add_inter = false, add_prev = false, add_cur = false;
    if (d2 == -1) {
      in_plane = false;
      if (d1 == 1) add_inter = true;
      else if (d1 == 0) add_prev = true;
    } else if (d2 == 1) {
      in_plane = false;
      if (d1 == -1) add_inter = true;
      else if (d1 == 0) add_prev = true;
      add_cur = true;
    } else if (d1 != 0) {
      add_cur = true;
    }
    if (add_inter && res.n < 16) {
      const double al = (cur[axis] - plane) / (cur[axis] - prev[axis]);
      for (int c = 0; c < 3; ++c) res.p[res.n][c] = al * prev[c] + (1.0 - al) * cur[c];
      ++res.n;
    }
    if (add_prev && res.n < 16) {
      const bool dup = res.n > 0 && res.p[res.n - 1][0] == prev[0] && res.p[res.n - 1][1] == prev[1] && res.p[res.n - 1][2] == prev[2];
      if (!dup) {
        for (int c = 0; c < 3; ++c) res.p[res.n][c] = prev[c];
        ++res.n;
      }
    }
    if (add_cur && res.n < 16) {
      for (int c = 0; c < 3; ++c) res.p[res.n][c] = cur[c];
      ++res.n;
    }
  }
  if (in_plane) return true;
  poly = res;
  return false;
}

static const int IOU_MAX_POINTS = 160;      // 2 x (6 faces x <= 10 clipped vertices + 8 inside vertices), with slack
struct Cloud {
  double p[IOU_MAX_POINTS][3];
  int n;
  int start[12], count[12], faces;           // the clipped face polygons inside p (the inside vertices are loose points)
  bool touched;                              // some vertex fell inside a thick clipping plane
};

TD3D_HD void cloud_add(Cloud& c, const double x[3]) {
  if (c.n >= IOU_MAX_POINTS) return;
  c.p[c.n][0] = x[0]; c.p[c.n][1] = x[1]; c.p[c.n][2] = x[2];
  ++c.n;
}

// iou.py:74-97: the faces of `tpl` clipped against `src` in src's local frame, and the vertices of `tpl` inside `src`,
// back in world space.  (The reference also adds tpl's centre point when it is inside: an interior point, no effect on the hull.)
TD3D_HD bool intersection_points(const Box& src, const Box& tpl, Cloud& cloud) {
  double inv[3][3];
  if (!inv3(src.r, inv)) return false;      // np.linalg.inv raises: the reference counts the pair as 0
  double rl[3][3], tl[3];                   // template in src's local frame: x_l = inv (R_t x + t_t - t_s)
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) rl[i][j] = inv[i][0] * tpl.r[0][j] + inv[i][1] * tpl.r[1][j] + inv[i][2] * tpl.r[2][j];
    tl[i] = inv[i][0] * (tpl.t[0] - src.t[0]) + inv[i][1] * (tpl.t[1] - src.t[1]) + inv[i][2] * (tpl.t[2] - src.t[2]);
  }
  const int faces[6][4] = {{5, 6, 8, 7}, {1, 3, 4, 2}, {3, 7, 8, 4}, {1, 2, 6, 5}, {2, 4, 8, 6}, {1, 5, 7, 3}};
  double vl[9][3];
  for (int i = 1; i < 9; ++i) {
    double u[3];
    unit_corner(i, u);
    for (int c = 0; c < 3; ++c) u[c] *= tpl.s[c];
    for (int c = 0; c < 3; ++c) vl[i][c] = rl[c][0] * u[0] + rl[c][1] * u[1] + rl[c][2] * u[2] + tl[c];
  }
  for (int f = 0; f < 6; ++f) {
    Poly poly;
    poly.n = 4;
    for (int k = 0; k < 4; ++k)
      for (int c = 0; c < 3; ++c) poly.p[k][c] = vl[faces[f][k]][c];
    for (int ax = 0; ax < 3; ++ax) {
      clip_poly(poly, -0.5 * src.s[ax], 1.0, ax, cloud.touched);
      clip_poly(poly, 0.5 * src.s[ax], -1.0, ax, cloud.touched);
    }
    cloud.start[cloud.faces] = cloud.n;
    for (int k = 0; k < poly.n; ++k) {
      double w[3];
      for (int c = 0; c < 3; ++c) w[c] = src.r[c][0] * poly.p[k][0] + src.r[c][1] * poly.p[k][1] + src.r[c][2] * poly.p[k][2] + src.t[c];
      cloud_add(cloud, w);
    }
    cloud.count[cloud.faces] = cloud.n - cloud.start[cloud.faces];
    ++cloud.faces;
  }
  for (int i = 1; i < 9; ++i) {
    if (fabs(vl[i][0]) > 0.5 * src.s[0] || fabs(vl[i][1]) > 0.5 * src.s[1] || fabs(vl[i][2]) > 0.5 * src.s[2]) continue;
    double w[3];
    for (int c = 0; c < 3; ++c) w[c] = src.r[c][0] * vl[i][0] + src.r[c][1] * vl[i][1] + src.r[c][2] * vl[i][2] + src.t[c];
    cloud_add(cloud, w);
  }
  return true;
}

// Volume of the convex hull of the intersection points (the reference calls Qhull).  The intersection of two boxes is a
// convex polyhedron whose faces are exactly the 12 clipped polygons, so in general position the volume is the sum of the
// pyramids they close over an interior point (fan triangles).  When a vertex fell inside a thick clipping plane (faces of
// the two boxes (nearly) coincide: the Sutherland-Hodgman pass then drops or duplicates polygons, which the hull of the point
// cloud does not mind) the hull itself is computed, by brute force: O(N^4) on N ~ 30 distinct points, only for such pairs.
TD3D_HD double hull_volume(const Cloud& cloud) {
  if (cloud.n < 4) return 0.0;
  double centre[3] = {0.0, 0.0, 0.0};
  for (int i = 0; i < cloud.n; ++i)
    for (int c = 0; c < 3; ++c) centre[c] += cloud.p[i][c];
  for (int c = 0; c < 3; ++c) centre[c] /= (double)cloud.n;
  if (!cloud.touched) {
    double vol6 = 0.0;
    for (int f = 0; f < cloud.faces; ++f) {
      if (cloud.count[f] < 3) continue;
      const double (*w)[3] = cloud.p + cloud.start[f];
      const double a0 = w[0][0] - centre[0], a1 = w[0][1] - centre[1], a2 = w[0][2] - centre[2];
      for (int k = 1; k + 1 < cloud.count[f]; ++k) {
        const double b0 = w[k][0] - centre[0], b1_ = w[k][1] - centre[1], b2_ = w[k][2] - centre[2];
        const double c0 = w[k + 1][0] - centre[0], c1 = w[k + 1][1] - centre[1], c2 = w[k + 1][2] - centre[2];
        vol6 += fabs(a0 * (b1_ * c2 - b2_ * c1) - a1 * (b0 * c2 - b2_ * c0) + a2 * (b0 * c1 - b1_ * c0));
      }
    }
    return vol6 / 6.0;
  }
  // exact hull of the point cloud, the way Qhull's joggle option finds it: drop repeated points, move every point by a
  // deterministic ~1e-9 so that no four are coplanar, then every triple with all other points on one side is a facet
  double q[IOU_MAX_POINTS][3];
  int n = 0;
  for (int i = 0; i < cloud.n; ++i) {
    bool dup = false;
    for (int j = 0; j < n && !dup; ++j)
      dup = fabs(q[j][0] - cloud.p[i][0]) < 1e-12 && fabs(q[j][1] - cloud.p[i][1]) < 1e-12 && fabs(q[j][2] - cloud.p[i][2]) < 1e-12;
    if (dup) continue;
    for (int c = 0; c < 3; ++c) {
      unsigned h = (unsigned)(3 * n + c + 1);        // integer hash (an affine sequence would move coplanar points coplanarly)
      h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
      q[n][c] = cloud.p[i][c] + ((double)(h >> 8) / 16777216.0 - 0.5) * 2e-9;
    }
    ++n;
  }
  if (n < 4) return 0.0;
  double vol6 = 0.0;
#pragma unroll 1
  for (int i = 0; i < n; ++i)
#pragma unroll 1
    for (int j = i + 1; j < n; ++j)
#pragma unroll 1
      for (int k = j + 1; k < n; ++k) {
        const double u0 = q[j][0] - q[i][0], u1 = q[j][1] - q[i][1], u2 = q[j][2] - q[i][2];
        const double v0 = q[k][0] - q[i][0], v1 = q[k][1] - q[i][1], v2 = q[k][2] - q[i][2];
        const double nx = u1 * v2 - u2 * v1, ny = u2 * v0 - u0 * v2, nz = u0 * v1 - u1 * v0;
        bool pos = false, neg = false;
        for (int r = 0; r < n && !(pos && neg); ++r) {
          if (r == i || r == j || r == k) continue;
          const double e = nx * (q[r][0] - q[i][0]) + ny * (q[r][1] - q[i][1]) + nz * (q[r][2] - q[i][2]);
          pos |= e > 0.0;
          neg |= e < 0.0;
        }
        if (pos && neg) continue;
        const double a0 = q[i][0] - centre[0], a1 = q[i][1] - centre[1], a2 = q[i][2] - centre[2];
        const double b0 = q[j][0] - centre[0], b1_ = q[j][1] - centre[1], b2_ = q[j][2] - centre[2];
        const double c0 = q[k][0] - centre[0], c1 = q[k][1] - centre[1], c2 = q[k][2] - centre[2];
        vol6 += fabs(a0 * (b1_ * c2 - b2_ * c1) - a1 * (b0 * c2 - b2_ * c0) + a2 * (b0 * c1 - b1_ * c0));
      }
  return vol6 / 6.0;
}

// iou.py:22-35
TD3D_HD double box_iou(const Box& b1, const Box& b2) {
  Cloud cloud;
  cloud.n = 0;
  cloud.faces = 0;
  cloud.touched = false;
  if (!intersection_points(b1, b2, cloud) || !intersection_points(b2, b1, cloud)) return 0.0;
  if (cloud.n == 0) return 0.0;
  const double inter = hull_volume(cloud);
  const double uni = b1.vol + b2.vol - inter;
  return uni > 0.0 ? inter / uni : 0.0;
}

// metrics.py:78-82 for one pair of keypoint sets
TD3D_HD double iou_from_keypoints(const float* pred, const float* gt, int portrait, const double cam[4]) {
  double lp[9][3], lg[9][3];
  lift_2d(pred, portrait, cam, lp);
  lift_2d(gt, portrait, cam, lg);
  Box bp, bg;
  box_fit(lp, bp);
  box_fit(lg, bg);
  return box_iou(bp, bg);
}

}  // namespace iou3d
}  // namespace td3d
