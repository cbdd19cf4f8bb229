// C++ host mirror smoke test: same calls a reference user makes (A.h:91-184), over the Mat stand-in.
// usage: test_compat <H> <W> <D> <in_L.bin> <in_R.bin> <out.bin>   (GPU)   |   test_compat --no-gpu
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "asw/aswMethods_compat.h"

using namespace asw_b200;

static Mat load_bgr(const char* path, int H, int W) {
    Mat m(H, W, 3, 1);
    FILE* f = fopen(path, "rb");
    if (!f || fread(m.data, 1, (size_t)H * W * 3, f) != (size_t)H * W * 3) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return m;
}

int main(int argc, char** argv) {
    if (argc >= 2 && strcmp(argv[1], "--no-gpu") == 0) {
        // without a device every method returns an empty Mat (no CPU fallback), even windows too
        Mat L(8, 8, 3, 1), R(8, 8, 3, 1);
        Mat d = computeAdaptiveWeight_GuidedF_2(L, R, DISPARITY_LEFT, 1e-4, 9, 0, 4);
        printf("empty=%d\n", d.empty() ? 1 : 0);
        return d.empty() ? 0 : 1;
    }
    if (argc < 7) return 2;
    int H = atoi(argv[1]), W = atoi(argv[2]), D = atoi(argv[3]);
    Mat L = load_bgr(argv[4], H, W), R = load_bgr(argv[5], H, W);
    FILE* f = fopen(argv[6], "wb");
    Mat disp;
    stereoMatching(L, R, disp, DISPARITY_LEFT, ADAPTIVE_WEIGHT_GUIDED_FILTER_2, 9, 0, D);     // main.cpp:94
    if (disp.empty()) return 3;
    fwrite(disp.data, 4, (size_t)H * W, f);
    Mat d2 = computeAdaptiveWeight(L, R, 30, 20, DISPARITY_LEFT, 7, 0, D);
    if (d2.empty()) return 4;
    fwrite(d2.data, 4, (size_t)H * W, f);
    Mat bad = computeAdaptiveWeight_geodesic(L, R, DISPARITY_LEFT, 8, 0, D);                    // even window -> Mat()
    if (!bad.empty()) return 5;
    Mat oos;
    stereoMatching(L, R, oos, DISPARITY_LEFT, SGBM, 9, 0, D);                                  // out of the hot path
    if (!oos.empty()) return 6;
    fclose(f);
    return 0;
}
