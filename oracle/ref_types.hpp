// Reduced stand-ins for Feature.hpp:105-143 (FeatureVec) and :152-158 (FeatureImage): only the
// members the scoring loops touch (fv, sqrt_simdot); fv is a pointer so D is not fixed at 342.
struct FeatureVec { int fid; float sqrt_simdot; const float* fv; };
typedef std::vector<FeatureVec> FVecV;
typedef std::map<Person, FVecV> P_FV;
struct FeatureImage { float* fv; float sqrt_simdot; };
typedef std::vector<FeatureImage> FVV;
