/* Plain-C restatement of the reference's native scoring path.  TEST INFRASTRUCTURE ONLY
 * (checker for tests/, smoke() and bench.py's cpu_baseline; never on the product path).
 * Follows feature_extraction/c_version/Feature.hpp of the reference:
 *   port_simd_dot          <- simd_dot                         :273-293
 *   port_gallery_scores    <- the sim expression               :346,356 / :776
 *   port_gallery_argmax    <- Compare_Face_From_DB(FVV&,..)    :345-392 (serial semantics)
 *   port_gallery_topk      <- Compare_Face_From_DB(Configs,..) :295-343 (list logic only)
 * No SSE intrinsics: the 4 accumulator lanes are written out so the rounding order is the
 * reference's (mul and add rounded separately; build with -ffp-contract=off).
 */
#include <math.h>
#include <stddef.h>

float port_simd_dot(const float* x, const float* y, long len) {
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  long i;
  for (i = 0; i + 4 < len; i += 4) {          /* strict '<': last full quad goes to the tail */
    a0 = a0 + x[i] * y[i];
    a1 = a1 + x[i + 1] * y[i + 1];
    a2 = a2 + x[i + 2] * y[i + 2];
    a3 = a3 + x[i + 3] * y[i + 3];
  }
  float s = a0 + a1 + a2 + a3;                /* ((t0+t1)+t2)+t3 */
  for (; i < len; ++i) s += x[i] * y[i];
  return s;
}

void port_gallery_scores(const float* G, const float* gnorm, long N, long D, const float* q, float* out) {
  float qn = sqrtf(port_simd_dot(q, q, D));
  long r;
#pragma omp parallel for schedule(static)
  for (r = 0; r < N; ++r) out[r] = port_simd_dot(G + (size_t)r * D, q, D) / (gnorm[r] * qn);
}

int port_gallery_argmax(const float* G, const float* gnorm, long N, long D, float sim_th, const float* q,
                        float* scratch, float* maxsim_out) {
  port_gallery_scores(G, gnorm, N, D, q, scratch);
  float maxsim = 0.f; int rt = -1; long r;
  for (r = 0; r < N; ++r) {
    float sim = scratch[r];
    if (isnan(sim)) continue;
    if (sim < sim_th) continue;
    if (sim > maxsim) { maxsim = sim; rt = (int)r; }
  }
  if (maxsim_out) *maxsim_out = maxsim;
  return rt;
}

/* top-K with the reference's replace-the-minimum-if-strictly-smaller rule; outputs sorted by
 * (sim desc, idx asc); unused slots get idx -1.  Returns -1 on a NaN similarity like :314-317. */
int port_gallery_topk(const float* G, const float* gnorm, long N, long D, float sim_th, const float* q,
                      int K, float* scratch, float* out_sim, int* out_idx) {
  port_gallery_scores(G, gnorm, N, D, q, scratch);
  int cnt = 0, k, status = 0; long r;
  for (k = 0; k < K; ++k) { out_sim[k] = 0.f; out_idx[k] = -1; }
  for (r = 0; r < N; ++r) {
    float sim = scratch[r];
    if (isnan(sim)) { status = -1; break; }
    if (sim < sim_th) continue;
    if (cnt < K) { out_sim[cnt] = sim; out_idx[cnt] = (int)r; ++cnt; }
    else {
      int m = 0;                       /* current minimum; among equal minima the latest row */
      for (k = 1; k < K; ++k)
        if (out_sim[k] < out_sim[m] || (out_sim[k] == out_sim[m] && out_idx[k] > out_idx[m])) m = k;
      if (out_sim[m] < sim) { out_sim[m] = sim; out_idx[m] = (int)r; }
    }
  }
  for (k = 1; k < cnt; ++k) {          /* insertion sort: sim desc, idx asc */
    float s = out_sim[k]; int id = out_idx[k]; int j = k - 1;
    while (j >= 0 && (out_sim[j] < s || (out_sim[j] == s && out_idx[j] > id))) {
      out_sim[j + 1] = out_sim[j]; out_idx[j + 1] = out_idx[j]; --j;
    }
    out_sim[j + 1] = s; out_idx[j + 1] = id;
  }
  return status;
}
