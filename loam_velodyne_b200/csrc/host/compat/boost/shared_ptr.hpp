// Stand-in used only when Boost is not installed: the reference's API types `pcl::PointCloud<T>::Ptr` as
// boost::shared_ptr (see include/loam_velodyne/nanoflann_pcl.h:44 upstream); std::shared_ptr has the same surface.
#pragma once
#include <memory>
namespace boost {
template <typename T>
using shared_ptr = std::shared_ptr<T>;
}
