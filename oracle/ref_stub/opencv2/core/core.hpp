// Stand-in for <opencv2/core/core.hpp>: just enough declarations for the reference's src/matrix_utils.cc to COMPILE here
// (OpenCV's C++ headers are not installed).  read_yaml() / bboxOverlapratio() are never called by oracle/_ref.
#ifndef PPO_STUB_OPENCV_CORE_HPP
#define PPO_STUB_OPENCV_CORE_HPP
#include <string>
#define CV_32F 5
namespace cv {
template <typename T>
struct Rect_ {
  T x, y, width, height;
  Rect_() : x(0), y(0), width(0), height(0) {}
  Rect_(T x_, T y_, T w_, T h_) : x(x_), y(y_), width(w_), height(h_) {}
  T area() const { return width * height; }
  Rect_ operator&(const Rect_ &o) const {
    const T x1 = x > o.x ? x : o.x, y1 = y > o.y ? y : o.y;
    const T x2 = (x + width < o.x + o.width) ? x + width : o.x + o.width, y2 = (y + height < o.y + o.height) ? y + height : o.y + o.height;
    return (x2 > x1 && y2 > y1) ? Rect_(x1, y1, x2 - x1, y2 - y1) : Rect_();
  }
  Rect_ operator|(const Rect_ &o) const {
    const T x1 = x < o.x ? x : o.x, y1 = y < o.y ? y : o.y;
    const T x2 = (x + width > o.x + o.width) ? x + width : o.x + o.width, y2 = (y + height > o.y + o.height) ? y + height : o.y + o.height;
    return Rect_(x1, y1, x2 - x1, y2 - y1);
  }
};
typedef Rect_<int> Rect;
class Mat {
  float v[16];
 public:
  int rows, cols;
  Mat() : rows(0), cols(0) {}
  Mat(int r, int c, int) : rows(r), cols(c) {}
  void resize(int r) { rows = r; }
  static Mat eye(int, int, int) { return Mat(); }
  template <typename T> T &at(int i) { return reinterpret_cast<T &>(v[i & 15]); }
  template <typename T> T &at(int i, int j) { return reinterpret_cast<T &>(v[(4 * i + j) & 15]); }
  void copyTo(Mat &o) const { o = *this; }
  Mat clone() const { return *this; }
};
struct FileNode {
  operator float() const { return 0.f; }
  operator double() const { return 0.0; }
  operator int() const { return 0; }
  operator std::string() const { return std::string(); }
};
class FileStorage {
 public:
  enum { READ = 0 };
  FileStorage(const std::string &, int) {}
  FileNode operator[](const char *) const { return FileNode(); }
  FileNode operator[](const std::string &) const { return FileNode(); }
  bool isOpened() const { return false; }
};
}  // namespace cv
#endif
