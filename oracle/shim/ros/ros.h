// TEST INFRASTRUCTURE (oracle/shim): the handful of roscpp names the reference's ROS adapters mention, so that
// src/lib/ScanRegistration.cpp and src/lib/MultiScanRegistration.cpp compile unmodified and the ROS-free body of
// MultiScanRegistration::process (ring binning, MultiScanRegistration.cpp:160-238) can be executed as the oracle.
// Nothing here talks to a ROS master: parameters are "not set", publishers swallow their messages.
#pragma once
#include <cstdint>
#include <string>

#include <boost/shared_ptr.hpp>

#define ROS_ERROR(...) ((void)0)
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)

namespace ros {

struct Time {
  uint32_t sec = 0, nsec = 0;
  Time& fromNSec(uint64_t t) {
    sec = (uint32_t)(t / 1000000000ull);
    nsec = (uint32_t)(t % 1000000000ull);
    return *this;
  }
};

inline bool operator==(const Time& a, const Time& b) { return a.sec == b.sec && a.nsec == b.nsec; }
inline bool operator!=(const Time& a, const Time& b) { return !(a == b); }
struct Duration { double s = 0; double toSec() const { return s; } };
inline Duration operator-(const Time& a, const Time& b) {
  Duration d;
  d.s = ((double)a.sec - (double)b.sec) + 1e-9 * ((double)a.nsec - (double)b.nsec);
  return d;
}
struct Rate {
  explicit Rate(double) {}
  void sleep() {}
};
inline bool ok() { return false; }  // no master: the adapters' spin() loops end at once
inline void spinOnce() {}

struct Publisher {
  template <typename M>
  void publish(const M&) const {}
};
struct Subscriber {};

class NodeHandle {
 public:
  template <typename T>
  bool getParam(const std::string&, T&) const { return false; }
  bool hasParam(const std::string&) const { return false; }
  template <typename M, typename T>
  Subscriber subscribe(const std::string&, uint32_t, void (T::*)(const boost::shared_ptr<M const>&), T*) { return Subscriber(); }
  template <typename M>
  Publisher advertise(const std::string&, uint32_t) { return Publisher(); }
};

}  // namespace ros
