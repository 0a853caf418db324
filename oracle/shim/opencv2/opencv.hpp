// TEST INFRASTRUCTURE ONLY (oracle build).  Minimal stand-in for the OpenCV 1.x C API
// symbols the reference's image writer touches (R/src/scene/film.cpp:39-64), so that the
// unmodified reference sources compile in an image without OpenCV development headers.
// cvSaveImage writes a binary PPM (BGR -> RGB), which is all the oracle harness needs.
#ifndef WRT_ORACLE_OPENCV_SHIM_HPP
#define WRT_ORACLE_OPENCV_SHIM_HPP
#include <cstdio>
#include <cstdlib>
typedef unsigned char uchar;
#define IPL_DEPTH_8U 8
struct CvSize { int width, height; };
struct IplImage { int width, height, nChannels, depth, widthStep; char* imageData; };
static inline CvSize cvSize(int w, int h) { CvSize s; s.width = w; s.height = h; return s; }
static inline IplImage* cvCreateImage(CvSize s, int depth, int ch) {
    IplImage* im = (IplImage*)std::calloc(1, sizeof(IplImage));
    im->width = s.width; im->height = s.height; im->nChannels = ch; im->depth = depth;
    im->widthStep = s.width * ch;
    im->imageData = (char*)std::calloc((size_t)im->widthStep * s.height, 1);
    return im;
}
static inline void cvReleaseImage(IplImage** im) {
    if (im && *im) { std::free((*im)->imageData); std::free(*im); *im = 0; }
}
static inline int cvSaveImage(const char* fn, const IplImage* im, const int* = 0) {
    FILE* f = std::fopen(fn, "wb");
    if (!f) return 0;
    std::fprintf(f, "P6\n%d %d\n255\n", im->width, im->height);
    for (int y = 0; y < im->height; y++)
        for (int x = 0; x < im->width; x++) {
            const uchar* p = (const uchar*)im->imageData + y * im->widthStep + x * im->nChannels;
            uchar rgb[3] = { p[2], p[1], p[0] };
            std::fwrite(rgb, 1, 3, f);
        }
    std::fclose(f);
    return 1;
}
static inline IplImage* cvLoadImage(const char*, int = 1) { return 0; }
#endif
