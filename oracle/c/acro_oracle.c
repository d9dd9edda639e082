/*
 * acro_oracle.c - plain-C restatement of the acrobotics FK + collision hot path
 * (TEST INFRASTRUCTURE: checker for 10^6..10^7-configuration parity runs, and the
 * native Box-Box call used by oracle/scalar_port.py in place of python-fcl).
 * Never linked into the product.
 *
 * Follows the reference (paths relative to /root/reference):
 *   oracle_dh            src/acrobotics/link.py:35-58
 *   oracle_fk_all        src/acrobotics/robot.py:82-91 (chain seeded with tf_base)
 *   oracle_box_box       python-fcl -> FCL boxBox2, reached from
 *                        src/acrobotics/shapes.py:50-58 (see oracle/fcl_boxbox.py)
 *   oracle_fk_cc         src/acrobotics/robot.py:244-273 + :206-221 +
 *                        src/acrobotics/geometry.py:51-66
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off acro_oracle.c -lm
 * (-ffp-contract=off: no fused multiply-add, like the numpy reference).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#define FUDGE2 1.0e-6

/* T (4x4 row-major) <- DH(a, alpha, d, theta) */
void oracle_dh(double a, double alpha, double d, double theta, double* T) {
  const double ct = cos(theta), st = sin(theta), ca = cos(alpha), sa = sin(alpha);
  memset(T, 0, 16 * sizeof(double));
  T[0] = ct;  T[1] = -st * ca; T[2] = st * sa;   T[3] = a * ct;
  T[4] = st;  T[5] = ct * ca;  T[6] = -ct * sa;  T[7] = a * st;
  T[9] = sa;  T[10] = ca;      T[11] = d;
  T[15] = 1.0;
}

static void mat4_mul(const double* A, const double* B, double* C) {
  double tmp[16];
  for (int r = 0; r < 4; ++r)
    for (int c = 0; c < 4; ++c) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += A[4 * r + k] * B[4 * k + c];
      tmp[4 * r + c] = s;
    }
  memcpy(C, tmp, sizeof(tmp));
}

/* frames: ndof 4x4 matrices */
void oracle_fk_all(int ndof, const double* dh /* ndof x 4: a, alpha, d, theta */,
                   const int32_t* prismatic, const double* tf_base, const double* q,
                   double* frames) {
  double T[16], Ti[16];
  memcpy(T, tf_base, sizeof(T));
  for (int i = 0; i < ndof; ++i) {
    const double* p = dh + 4 * i;
    if (prismatic[i])
      oracle_dh(p[0], p[1], q[i], p[3], Ti);
    else
      oracle_dh(p[0], p[1], p[2], q[i], Ti);
    mat4_mul(T, Ti, T);
    memcpy(frames + 16 * i, T, sizeof(T));
  }
}

/* FCL boxBox2 decision.  side*: full side lengths; tf*: 4x4 row-major.
 * Returns 1 when the boxes collide; *g_out (nullable) = largest decision quantity. */
int oracle_box_box(const double* side1, const double* tf1, const double* side2, const double* tf2,
                   double* g_out) {
  double p[3], pp[3], A[3], B[3], R[9], Q[9];
  for (int r = 0; r < 3; ++r) {
    p[r] = tf2[4 * r + 3] - tf1[4 * r + 3];
    A[r] = 0.5 * side1[r];
    B[r] = 0.5 * side2[r];
  }
  for (int i = 0; i < 3; ++i) {
    pp[i] = tf1[i] * p[0] + tf1[4 + i] * p[1] + tf1[8 + i] * p[2];
    for (int j = 0; j < 3; ++j) {
      R[3 * i + j] = tf1[i] * tf2[j] + tf1[4 + i] * tf2[4 + j] + tf1[8 + i] * tf2[8 + j];
      Q[3 * i + j] = fabs(R[3 * i + j]);
    }
  }
  double g = -DBL_MAX;
  int collide = 1;
  for (int i = 0; i < 3; ++i) {
    const double s2 = fabs(pp[i]) - (Q[3 * i] * B[0] + Q[3 * i + 1] * B[1] + Q[3 * i + 2] * B[2] + A[i]);
    if (s2 > g) g = s2;
    if (s2 > 0) collide = 0;
  }
  for (int j = 0; j < 3; ++j) {
    const double t = tf2[j] * p[0] + tf2[4 + j] * p[1] + tf2[8 + j] * p[2];
    const double s2 = fabs(t) - (Q[j] * A[0] + Q[3 + j] * A[1] + Q[6 + j] * A[2] + B[j]);
    if (s2 > g) g = s2;
    if (s2 > 0) collide = 0;
  }
  for (int k = 0; k < 9; ++k) Q[k] += FUDGE2;
  for (int i = 0; i < 3; ++i) {
    const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
    for (int j = 0; j < 3; ++j) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      const double t = pp[i2] * R[3 * i1 + j] - pp[i1] * R[3 * i2 + j];
      const double s2 = fabs(t) - (A[i1] * Q[3 * i2 + j] + A[i2] * Q[3 * i1 + j] +
                                   B[j1] * Q[3 * i + j2] + B[j2] * Q[3 * i + j1]);
      if (s2 - DBL_EPSILON > g) g = s2 - DBL_EPSILON;
      if (s2 > DBL_EPSILON) collide = 0;
    }
  }
  if (g_out) *g_out = g;
  return collide;
}

/*
 * Robot.is_in_collision for n configurations.
 *   boxes: n_boxes x (carrier, 3 sides, 16 tf) packed as 20 doubles each; carrier
 *          >= 0 link index, -1 base, -2 tool (last link frame)
 *   pairs: n_pairs x 2 box indices (self collision)
 *   scene: n_scene x (3 sides, 16 tf) = 19 doubles each
 * gmin (nullable, n): smallest pair-wise largest decision quantity.
 */
void oracle_fk_cc(int ndof, const double* dh, const int32_t* prismatic, const double* tf_base,
                  int n_boxes, const double* boxes, int n_pairs, const int32_t* pairs, int n_scene,
                  const double* scene, int check_self, const double* q, int64_t n, uint8_t* hit,
                  double* gmin) {
  double frames[8 * 16];
  double world[32 * 16];
  for (int64_t c = 0; c < n; ++c) {
    oracle_fk_all(ndof, dh, prismatic, tf_base, q + c * ndof, frames);
    for (int b = 0; b < n_boxes; ++b) {
      const double* bx = boxes + 20 * b;
      const int carrier = (int)bx[0];
      const double* F = carrier == -1 ? tf_base : frames + 16 * (carrier == -2 ? ndof - 1 : carrier);
      mat4_mul(F, bx + 4, world + 16 * b);
    }
    double g = DBL_MAX, gp;
    for (int b = 0; b < n_boxes; ++b)
      for (int s = 0; s < n_scene; ++s) {
        oracle_box_box(boxes + 20 * b + 1, world + 16 * b, scene + 19 * s, scene + 19 * s + 3, &gp);
        if (gp < g) g = gp;
      }
    if (check_self)
      for (int k = 0; k < n_pairs; ++k) {
        const int b1 = pairs[2 * k], b2 = pairs[2 * k + 1];
        oracle_box_box(boxes + 20 * b1 + 1, world + 16 * b1, boxes + 20 * b2 + 1, world + 16 * b2, &gp);
        if (gp < g) g = gp;
      }
    hit[c] = !(g > 0);
    if (gmin) gmin[c] = g;
  }
}
