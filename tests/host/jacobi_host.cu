// Host harness for the __host__ __device__ math in csrc/jacobi.cuh (tests/test_host_math.py).
// stdin-free: argv[1] = mode (dlt|dltbwd|pinv), argv[2] = input file, argv[3] = output file.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../fast-3d-human-pose-estimation_b200/csrc/jacobi.cuh"

int main(int argc, char** argv) {
  if (argc != 4) return 2;
  FILE* f = fopen(argv[2], "rb");
  if (!f) return 3;
  long long n;
  if (fread(&n, 8, 1, f) != 1) return 4;
  FILE* o = fopen(argv[3], "wb");
  if (!strcmp(argv[1], "dlt")) {           // per item: P_l[12] P_r[12] float, kp[4] float
    std::vector<float> in(n * 28);
    if (fread(in.data(), 4, in.size(), f) != in.size()) return 5;
    std::vector<double> out(n * 3);
    for (long long i = 0; i < n; ++i) {
      const float* p = &in[i * 28];
      double A[4][4];
      cdr::dlt_rows(p, (double)p[24], (double)p[25], A, 0);
      cdr::dlt_rows(p + 12, (double)p[26], (double)p[27], A, 2);
      cdr::dlt_solve4(A, out[i * 3], out[i * 3 + 1], out[i * 3 + 2]);
    }
    fwrite(out.data(), 8, out.size(), o);
  } else if (!strcmp(argv[1], "dltbwd")) {   // per item: P_l[12] P_r[12] kp[4] grad_xyz[3] float -> grad_kp[4] double
    std::vector<float> in(n * 31);
    if (fread(in.data(), 4, in.size(), f) != in.size()) return 5;
    std::vector<double> out(n * 4);
    for (long long i = 0; i < n; ++i) {
      const float* p = &in[i * 31];
      const double gX[3] = {p[28], p[29], p[30]};
      double g[4];
      cdr::dlt_backward4(p, p + 12, (double)p[24], (double)p[25], (double)p[26], (double)p[27], gX, g);
      for (int c = 0; c < 4; ++c) out[i * 4 + c] = g[c];
    }
    fwrite(out.data(), 8, out.size(), o);
  } else if (!strcmp(argv[1], "pinv")) {   // per item: P[12] float ; rtol double first
    double rtol;
    if (fread(&rtol, 8, 1, f) != 1) return 5;
    std::vector<float> in(n * 12), out(n * 12);
    if (fread(in.data(), 4, in.size(), f) != in.size()) return 5;
    for (long long i = 0; i < n; ++i) cdr::pinv_3x4<float, float>(&in[i * 12], rtol, &out[i * 12]);
    fwrite(out.data(), 4, out.size(), o);
  } else {
    return 6;
  }
  fclose(o);
  fclose(f);
#ifdef CDR_JACOBI_STATS
  fprintf(stderr, "avg sweeps per item: %.2f\n", (double)cdr_jacobi_sweeps / (double)n);
#endif
  return 0;
}
