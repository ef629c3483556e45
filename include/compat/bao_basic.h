// Minimal stand-in for the reference's basic/bao_basic.h: the caller-side memory convention of the class
// (basic/bao_basic.h:105-174) and the PPM loader main.cpp uses.  Layout contract: a 3-D array img[y][x][c] is ONE
// contiguous block reachable as img[0][0]; a 2-D array u[y][x] is one contiguous block reachable as u[0].
#ifndef _EPPM_COMPAT_BAO_BASIC_H_
#define _EPPM_COMPAT_BAO_BASIC_H_
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <iostream>
using std::cout;
using std::endl;

template <typename T> inline T* bao_alloc(int n) {
    T* p = static_cast<T*>(malloc(sizeof(T) * (size_t)n));
    if (!p) { fprintf(stderr, "bao_alloc: out of memory\n"); exit(1); }
    return p;
}
template <typename T> inline T** bao_alloc(int rows, int cols) {
    T* data = bao_alloc<T>(rows * cols);
    T** rp = static_cast<T**>(malloc(sizeof(T*) * (size_t)rows));
    for (int y = 0; y < rows; ++y) rp[y] = data + (size_t)y * cols;
    return rp;
}
template <typename T> inline T*** bao_alloc(int n, int rows, int cols) {
    T** rp = bao_alloc<T>(n * rows, cols);  // contiguous n*rows*cols block with one row pointer per (n,row)
    T*** pp = static_cast<T***>(malloc(sizeof(T**) * (size_t)n));
    for (int i = 0; i < n; ++i) pp[i] = rp + (size_t)i * rows;
    return pp;
}
template <typename T> inline void bao_free(T*& p) { free(p); p = NULL; }
template <typename T> inline void bao_free(T**& p) { if (p) { free(p[0]); free(p); p = NULL; } }
template <typename T> inline void bao_free(T***& p) { if (p) { free(p[0][0]); free(p[0]); free(p); p = NULL; } }

// Binary PGM/PPM (P5/P6) and ASCII (P2/P3) reader with '#' comment lines; fills h*w*channels bytes.
inline int bao_loadimage_ppm(const char* filename, unsigned char* image, int h, int w, int* nr_channel) {
    FILE* f = fopen(filename, "rb");
    if (!f) { printf("Please check input filename: %s\n", filename); exit(0); }
    auto token = [&](char* buf, int cap) {  // next whitespace-delimited token, skipping comments
        int c = fgetc(f), n = 0;
        for (;;) {
            while (c == ' ' || c == '\t' || c == '\r' || c == '\n') c = fgetc(f);
            if (c != '#') break;
            while (c != '\n' && c != EOF) c = fgetc(f);
        }
        while (c != EOF && c != ' ' && c != '\t' && c != '\r' && c != '\n' && n < cap - 1) { buf[n++] = (char)c; c = fgetc(f); }
        buf[n] = 0;
    };
    char t[64];
    token(t, sizeof t);
    if (t[0] != 'P') { printf("Bad header in ppm file.\n"); exit(1); }
    const int kind = atoi(t + 1);
    token(t, sizeof t); const int wc = atoi(t);
    token(t, sizeof t); const int hc = atoi(t);
    token(t, sizeof t);  // maxval; exactly one whitespace byte was consumed after it
    (void)wc; (void)hc;
    const int ch = (kind == 6 || kind == 3) ? 3 : 1;
    if (nr_channel) *nr_channel = ch;
    const size_t n = (size_t)h * w * ch;
    memset(image, 0, n);
    if (kind == 5 || kind == 6) {
        if (fread(image, 1, n, f) != n) printf("Short read in %s\n", filename);
    } else if (kind == 2 || kind == 3) {
        for (size_t i = 0; i < n; ++i) { int v = 0; if (fscanf(f, "%d", &v) != 1) break; image[i] = (unsigned char)v; }
    } else {
        printf("Can not open image [%s]!!\n", filename);
    }
    fclose(f);
    return 0;
}
#endif
