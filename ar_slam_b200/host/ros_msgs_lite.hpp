// ros_msgs_lite.hpp -- plain C++ stand-ins for the ROS2 message types the solver PRODUCES
// (geometry_msgs/TransformStamped, sensor_msgs/CameraInfo, visualization_msgs/Marker,
// builtin_interfaces/Time), used when the host is built without ROS2 (this repository's CI).
// Field names, types and constants follow the ROS2 Iron message definitions the reference
// fills at ar_slam/src/ar_slam_util.cpp:1027-1162; with ARSLAM_WITH_ROS the generated headers
// and rclcpp::Time are used instead, and ArSlamSolver's getters compile against them unchanged.
#pragma once
#ifdef ARSLAM_WITH_ROS
#include "geometry_msgs/msg/transform_stamped.hpp"
#include "rclcpp/time.hpp"
#include "sensor_msgs/distortion_models.hpp"
#include "sensor_msgs/msg/camera_info.hpp"
#include "visualization_msgs/msg/marker.hpp"
namespace arslam_ros { using Time = rclcpp::Time; }
#else
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace builtin_interfaces { namespace msg {
struct Time { int32_t sec = 0; uint32_t nanosec = 0; };
} }
namespace std_msgs { namespace msg {
struct Header { builtin_interfaces::msg::Time stamp; std::string frame_id; };
struct ColorRGBA { float r = 0.f, g = 0.f, b = 0.f, a = 0.f; };
} }
namespace geometry_msgs { namespace msg {
struct Vector3 { double x = 0.0, y = 0.0, z = 0.0; };
struct Quaternion { double x = 0.0, y = 0.0, z = 0.0, w = 1.0; };
struct Transform { Vector3 translation; Quaternion rotation; };
struct TransformStamped { std_msgs::msg::Header header; std::string child_frame_id; Transform transform; };
} }
namespace sensor_msgs {
namespace distortion_models { const std::string PLUMB_BOB = "plumb_bob"; }
namespace msg {
struct CameraInfo {
  std_msgs::msg::Header header;
  uint32_t height = 0, width = 0;
  std::string distortion_model;
  std::vector<double> d;
  std::array<double, 9> k{};
  std::array<double, 9> r{};
  std::array<double, 12> p{};
};
} }
namespace visualization_msgs { namespace msg {
struct Marker {
  static constexpr int32_t CUBE = 1;
  static constexpr int32_t ADD = 0, DELETEALL = 3;
  std_msgs::msg::Header header;
  std::string ns;
  int32_t id = 0, type = 0, action = 0;
  geometry_msgs::msg::Vector3 scale;
  std_msgs::msg::ColorRGBA color;
  bool frame_locked = false;
};
} }
namespace arslam_ros { using Time = builtin_interfaces::msg::Time; }
#endif
