// C entry points over the reference functions compiled above.  TEST INFRASTRUCTURE ONLY.
extern "C" {
void ref_set_fvsize(int d) { fvSize = d; }
float ref_simd_dot(const float* x, const float* y, long len) { return simd_dot(x, y, len); }
// Compare_Face_From_DB(FVV&, float, float*)  Feature.hpp:345
int ref_compare_face_from_db(const float* G, const float* gnorm, int N, int D, float sim_th, const float* q) {
  fvSize = D;
  FVV mbs(N);
  for (int i = 0; i < N; ++i) { mbs[i].fv = const_cast<float*>(G + (size_t)i * D); mbs[i].sqrt_simdot = gnorm[i]; }
  return Compare_Face_From_DB(mbs, sim_th, const_cast<float*>(q));
}
// Compare_Face_DB(P_FV&, float, float*)  Feature.hpp:792 ; returns pid (-1 when nobody matches)
int ref_compare_face_db(const float* G, const float* gnorm, const int* pids, int N, int D, float sim_th, const float* q) {
  fvSize = D;
  P_FV mbs;
  for (int i = 0; i < N; ++i) {
    Person p; p.pid = pids[i];
    FeatureVec f; f.fid = i; f.sqrt_simdot = gnorm[i]; f.fv = G + (size_t)i * D;
    mbs[p].push_back(f);
  }
  Person r = Compare_Face_DB(mbs, sim_th, const_cast<float*>(q));
  return r.pid;
}
// Compare_Face_Person(const FVecV&, float, float*, float&, bool, float)  Feature.hpp:763
int ref_compare_face_person(const float* G, const float* gnorm, int N, int D, float sim_th, const float* q,
                            float* maxsim, int is_cal_dot, float cal_dot) {
  fvSize = D;
  FVecV v(N);
  for (int i = 0; i < N; ++i) { v[i].fid = i; v[i].sqrt_simdot = gnorm[i]; v[i].fv = G + (size_t)i * D; }
  return Compare_Face_Person(v, sim_th, const_cast<float*>(q), *maxsim, is_cal_dot != 0, cal_dot) ? 1 : 0;
}
}
