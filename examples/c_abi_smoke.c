/* c_abi_smoke.c - libacro_b200.so from plain C: no Python, no torch.
 *
 * Builds the reference's Kuka (src/acrobotics/robot_examples.py:137-186) and the README scene
 * (README.md:72-79) by hand, checks the README path (README.md:84-96) on the GPU and prints
 * the verdicts.  Exit code 0 = the expected [F F F F T T T T F F].
 *
 *   gcc -std=c99 -I include -I /usr/local/cuda/include examples/c_abi_smoke.c \
 *       -L acrobotics_b200 -lacro_b200 -L /usr/local/cuda/lib64 -lcudart -lm -o c_abi_smoke
 */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include "acro_b200.h"

static void identity(double* tf) {
  memset(tf, 0, 12 * sizeof(double));
  tf[0] = tf[5] = tf[10] = 1.0;
}

int main(void) {
  const double pi = 3.14159265358979323846;
  const double a[6] = {0.18, 0.6, 0, 0, 0, 0};
  const double alpha[6] = {pi / 2, 0, pi / 2, -pi / 2, pi / 2, 0};
  const double d[6] = {0, 0, 0, 0.62, 0, 0.115};
  const double sides[6][3] = {{0.3, 0.2, 0.1},  {0.8, 0.2, 0.1},   {0.2, 0.1, 0.5},
                              {0.1, 0.2, 0.1},  {0.1, 0.1, 0.085}, {0.1, 0.1, 0.03}};
  const double offs[6][3] = {{-0.09, 0, 0.05}, {-0.3, 0, -0.05}, {0, 0.05, 0.17},
                             {0, 0.1, 0},      {0, 0, 0.0425},   {0, 0, -0.015}};
  static AcroRobotDesc desc;
  memset(&desc, 0, sizeof(desc));
  desc.ndof = 6;
  for (int i = 0; i < 6; ++i) {
    desc.a[i] = a[i];
    desc.d[i] = d[i];
    desc.cos_alpha[i] = cos(alpha[i]);
    desc.sin_alpha[i] = sin(alpha[i]);
    desc.boxes[i].carrier = i;
    for (int k = 0; k < 3; ++k) desc.boxes[i].half[k] = 0.5 * sides[i][k];
    identity(desc.boxes[i].tf);
    desc.boxes[i].tf[3] = offs[i][0];
    desc.boxes[i].tf[7] = offs[i][1];
    desc.boxes[i].tf[11] = offs[i][2];
  }
  identity(desc.tf_base);
  identity(desc.tf_tool_tip);
  desc.n_boxes = 6;
  desc.check_self = 1;
  /* collision_matrix[i][j] = |i - j| >= 3  (robot.py:185-186) */
  for (int i = 0; i < 6; ++i)
    for (int j = i + 3; j < 6; ++j) {
      desc.self_pairs[desc.n_self_pairs][0] = i;
      desc.self_pairs[desc.n_self_pairs][1] = j;
      ++desc.n_self_pairs;
    }

  AcroBox scene_boxes[2];
  memset(scene_boxes, 0, sizeof(scene_boxes));
  const double s0[3] = {2, 2, 0.1}, s1[3] = {0.2, 0.2, 1.5};
  for (int k = 0; k < 3; ++k) {
    scene_boxes[0].half[k] = 0.5 * s0[k];
    scene_boxes[1].half[k] = 0.5 * s1[k];
  }
  identity(scene_boxes[0].tf);
  identity(scene_boxes[1].tf);
  scene_boxes[0].tf[11] = -0.2;
  scene_boxes[1].tf[7] = 0.5;
  scene_boxes[1].tf[11] = 0.55;

  AcroRobot* robot = NULL;
  AcroScene* scene = NULL;
  if (acro_robot_create(&desc, &robot) || acro_scene_create(scene_boxes, 2, &scene)) {
    fprintf(stderr, "create failed: %s\n", acro_last_error());
    return 2;
  }

  double q[10][6];
  for (int i = 0; i < 10; ++i) { /* np.linspace([0.5,1.5,-0.3,0,0,0],[2.5,1.5,0.3,0,0,0],10) */
    const double s = i / 9.0;
    q[i][0] = 0.5 + 2.0 * s;
    q[i][1] = 1.5;
    q[i][2] = -0.3 + 0.6 * s;
    q[i][3] = q[i][4] = q[i][5] = 0.0;
  }
  double* d_q = NULL;
  uint32_t* d_mask = NULL;
  uint32_t mask = 0;
  if (cudaMalloc((void**)&d_q, sizeof(q)) || cudaMalloc((void**)&d_mask, sizeof(uint32_t)) ||
      cudaMemcpy(d_q, q, sizeof(q), cudaMemcpyHostToDevice)) {
    fprintf(stderr, "CUDA setup failed\n");
    return 3;
  }
  int rc = acro_fk_cc_f64(robot, scene, d_q, 10, ACRO_Q_AOS, d_mask, NULL, 0.0, NULL);
  if (rc || cudaMemcpy(&mask, d_mask, sizeof(mask), cudaMemcpyDeviceToHost)) {
    fprintf(stderr, "acro_fk_cc_f64 failed (%d): %s\n", rc, acro_last_error());
    return 4;
  }
  printf("verdicts:");
  for (int i = 0; i < 10; ++i) printf(" %c", (mask >> i) & 1u ? 'T' : 'F');
  printf("\n");
  cudaFree(d_q);
  cudaFree(d_mask);
  acro_scene_destroy(scene);
  acro_robot_destroy(robot);
  return mask == 0xF0u ? 0 : 1;
}
