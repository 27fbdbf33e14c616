// oracle/shim/opencv2 — NOT OpenCV: empty stand-ins for the drawing calls in the demos' main()s so that
// the reference sources compile where OpenCV is absent.  Nothing here is ever executed by the tests.
#ifndef CRB_SHIM_OPENCV_
#define CRB_SHIM_OPENCV_
#include <string>
namespace cv {
struct Point2i { int x, y; Point2i() : x(0), y(0) {} Point2i(int a, int b) : x(a), y(b) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Size { int w, h; Size(double a, double b) : w((int)a), h((int)b) {} };
struct Mat { int rows, cols; Mat() : rows(0), cols(0) {} Mat(int r, int c, int, Scalar = Scalar()) : rows(r), cols(c) {} };
enum { CV_8UC3_ = 16, WINDOW_NORMAL = 0, MARKER_CROSS = 0, FONT_HERSHEY_SIMPLEX = 0 };
inline void putText(Mat, const std::string&, Point2i, int, double, Scalar, int = 1) {}
inline void namedWindow(const std::string&, int = 0) {}
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return 0; }
inline bool imwrite(const std::string&, const Mat&) { return true; }
inline void circle(Mat, Point2i, int, Scalar, int = 1) {}
inline void line(Mat, Point2i, Point2i, Scalar, int = 1) {}
inline void ellipse(Mat, Point2i, Size, double, double, double, Scalar, int = 1, int = 8) {}
inline void drawMarker(Mat, Point2i, Scalar, int = 0, int = 20, int = 1) {}
inline void arrowedLine(Mat, Point2i, Point2i, Scalar, int = 1) {}
}  // namespace cv
#ifndef CV_8UC3
#define CV_8UC3 16
#endif
#endif
