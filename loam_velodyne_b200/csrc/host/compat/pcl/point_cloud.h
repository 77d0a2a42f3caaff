// Minimal pcl::PointCloud<T> for builds without PCL: exactly the members the Basic* classes and their ROS adapters
// use (points, push_back, +=, clear, size, [], iteration, is_dense, width/height, Ptr).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>
#include <pcl/point_types.h>

namespace pcl {

struct PCLHeader {
  std::uint32_t seq;
  std::uint64_t stamp;
  std::string frame_id;
  PCLHeader() : seq(0), stamp(0) {}
};

template <typename PointT>
class PointCloud {
 public:
  typedef boost::shared_ptr<PointCloud<PointT> > Ptr;
  typedef boost::shared_ptr<const PointCloud<PointT> > ConstPtr;
  typedef std::vector<PointT> VectorType;
  typedef typename VectorType::iterator iterator;
  typedef typename VectorType::const_iterator const_iterator;

  PointCloud() : width(0), height(0), is_dense(true) {}
  PointCloud(std::uint32_t w, std::uint32_t h) : points(static_cast<std::size_t>(w) * h), width(w), height(h), is_dense(true) {}

  PointCloud& operator+=(const PointCloud& rhs) {
    points.insert(points.end(), rhs.points.begin(), rhs.points.end());
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
    if (!rhs.is_dense) is_dense = false;
    return *this;
  }
  void push_back(const PointT& p) {
    points.push_back(p);
    width = static_cast<std::uint32_t>(points.size());
    height = 1;
  }
  void clear() { points.clear(); width = 0; height = 0; }
  void resize(std::size_t n) { points.resize(n); width = static_cast<std::uint32_t>(n); height = 1; }
  void reserve(std::size_t n) { points.reserve(n); }
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
  PointT& at(std::size_t i) { return points.at(i); }
  const PointT& at(std::size_t i) const { return points.at(i); }
  iterator begin() { return points.begin(); }
  iterator end() { return points.end(); }
  const_iterator begin() const { return points.begin(); }
  const_iterator end() const { return points.end(); }

  PCLHeader header;
  VectorType points;
  std::uint32_t width, height;
  bool is_dense;
};

}  // namespace pcl
