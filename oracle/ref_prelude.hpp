// Stand-in prelude for compiling the reference's own scoring functions
// (feature_extraction/c_version/Feature.hpp) out of tree.  TEST INFRASTRUCTURE ONLY.
// The reference header as a whole cannot compile here (needs mxnet/c_predict_api.h, OpenCV,
// sqlite3, boost and three headers missing from the reference repo), so build_ref.sh feeds the
// compiler: this prelude + Feature.hpp:33-103 (class Person, verbatim from the reference tree)
// + ref_types.hpp (reduced carriers for FeatureVec / FeatureImage without cv::Mat) +
// Feature.hpp:273-293 (simd_dot) + :345-392 (Compare_Face_From_DB in-memory) + :763-804
// (Compare_Face_Person, Compare_Face_DB) + ref_exports.cpp.  Nothing from the reference is copied
// into this repository; the lines are read from /root/reference at build time.
#include <xmmintrin.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <map>
#include <string>
#include <vector>
// Feature.hpp:24 has `const int fvSize = 342;` -- made settable so tests can use other D.
static int fvSize = 342;
