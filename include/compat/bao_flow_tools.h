// Minimal stand-in for the reference's basic/bao_flow_tools.h: Middlebury .flo writer/reader
// (format: "PIEH", int32 width, int32 height, then rows of interleaved (u,v) float32 — 3rdparty/middlebury/README.txt:9-24)
// and the end-point-error measure of bao_calc_flow_error (basic/bao_flow_tools.cpp:64-111).
#ifndef _EPPM_COMPAT_BAO_FLOW_TOOLS_H_
#define _EPPM_COMPAT_BAO_FLOW_TOOLS_H_
#include <stdio.h>
#include <vector>

inline void bao_save_flo_file(const char* filename, float** disp_x, float** disp_y, int h, int w) {
    FILE* f = fopen(filename, "wb");
    if (!f) { printf("bao_save_flo_file: could not open %s\n", filename); return; }
    fwrite("PIEH", 1, 4, f);
    fwrite(&w, sizeof(int), 1, f);
    fwrite(&h, sizeof(int), 1, f);
    std::vector<float> row((size_t)2 * w);
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) { row[2 * x] = disp_x[y][x]; row[2 * x + 1] = disp_y[y][x]; }
        fwrite(row.data(), sizeof(float), row.size(), f);
    }
    fclose(f);
}
inline bool bao_load_flo_file(const char* filename, float** disp_x, float** disp_y, int h, int w) {
    FILE* f = fopen(filename, "rb");
    if (!f) return false;
    char tag[4]; int fw = 0, fh = 0;
    bool ok = fread(tag, 1, 4, f) == 4 && fread(&fw, 4, 1, f) == 1 && fread(&fh, 4, 1, f) == 1 && fw == w && fh == h;
    std::vector<float> row((size_t)2 * w);
    for (int y = 0; ok && y < h; ++y) {
        ok = fread(row.data(), sizeof(float), row.size(), f) == row.size();
        for (int x = 0; ok && x < w; ++x) { disp_x[y][x] = row[2 * x]; disp_y[y][x] = row[2 * x + 1]; }
    }
    fclose(f);
    return ok;
}
#endif
