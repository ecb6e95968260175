// detections_msg.hpp -- plain C++ stand-ins for the ROS2 message types the solver
// consumes, used when the host is built without ROS2 (this repository's CI).
// Field names and types follow ar_slam_interfaces/msg/Detection.msg:1-2 and
// Detections.msg:1-19 of the reference; with ARSLAM_WITH_ROS the generated
// headers are used instead and this file is empty.
#pragma once
#ifdef ARSLAM_WITH_ROS
#include "ar_slam_interfaces/msg/detections.hpp"
#else
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace geometry_msgs { namespace msg {
struct Point32 { float x = 0.f, y = 0.f, z = 0.f; };
} }

namespace ar_slam_interfaces { namespace msg {
struct Detection {
  std::array<geometry_msgs::msg::Point32, 4> corners;  // centred pixels, TL,TR,BR,BL (float32!)
  std::string id;                                      // e.g. aruco_4X4_50_18
};
struct Detections {
  std::string capture_uid;
  uint32_t image_height = 0;
  uint32_t image_width = 0;
  std::string image_path;
  std::vector<std::string> detector_types;
  std::vector<Detection> detections;
};
} }
#endif
