// TEST INFRASTRUCTURE (oracle shim): boost::shared_ptr -> std::shared_ptr (Boost is absent from this image;
// the reference only uses it as the PCL cloud handle, nanoflann_pcl.h:44, pcl::PointCloud<T>::Ptr).
#pragma once
#include <memory>
namespace boost {
template <typename T> using shared_ptr = std::shared_ptr<T>;
}
