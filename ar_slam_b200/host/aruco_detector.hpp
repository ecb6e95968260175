// aruco_detector.hpp -- host-side drop-in for the reference's detector node without rclcpp / OpenCV.
//
// Mirrors ar_slam::ArucoDetector (reference ar_slam/src/aruco_detector.cpp:40-152): a dictionary chosen by name
// ("4X4_50" default, "5X5_100", "6X6_250"; anything else throws "invalid aruco_dict", :71-76), and per image the
// body of image_callback (:95-140): detectMarkers, then a Detections message with detector_types =
// {"aruco_<dict>"}, ids "aruco_<dict>_<n>" and the four corners centred on the image (from_cv_img,
// ar_slam_util.hpp:257-263) as float32 Point32.  cv::aruco::detectMarkers is replaced by the CUDA library's
// arslam_detect_markers (include/ar_slam_b200.h); frames arrive decoded (8-bit BGR or grey), several equally
// sized frames go through the GPU in one call.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/ar_slam_b200.h"
#include "detections_msg.hpp"

namespace ar_slam {

struct Frame {                 // what reaches image_callback as ar_slam_interfaces::msg::Capture
  const uint8_t* data = nullptr;   // height x width x channels, tightly packed; 3 channels = BGR (cv_bridge / cv::imread)
  int width = 0, height = 0, channels = 3;
  std::string capture_uid, image_path;
};

class ArucoDetector {
public:
  explicit ArucoDetector(const std::string& dict_name = "4X4_50", int device = 0, int max_images = 16,
                         int max_width = 4096, int max_height = 4096)
      : detector_name_("aruco_" + dict_name) {
    arslam_detect_default_params(&params_);                 // cv::aruco::DetectorParameters::create() (:57)
    if (arslam_detector_create(device, max_images, max_width, max_height, &det_) != ARSLAM_OK)
      throw std::runtime_error(std::string("arslam_detector_create: ") + arslam_detector_last_error(nullptr));
    if (arslam_detector_set_predefined_dictionary(det_, dict_name.c_str()) != ARSLAM_OK) {
      arslam_detector_destroy(det_);
      throw std::runtime_error("invalid aruco_dict");       // :75
    }
  }
  ~ArucoDetector() { arslam_detector_destroy(det_); }
  ArucoDetector(const ArucoDetector&) = delete;
  ArucoDetector& operator=(const ArucoDetector&) = delete;

  arslam_detect_params& params() { return params_; }
  const std::string& detector_name() const { return detector_name_; }

  // image_callback (:95-140) for a batch of equally sized frames
  std::vector<ar_slam_interfaces::msg::Detections> detect(const std::vector<Frame>& frames) {
    std::vector<ar_slam_interfaces::msg::Detections> out(frames.size());
    if (frames.empty()) return out;
    const int w = frames[0].width, h = frames[0].height, ch = frames[0].channels;
    const size_t bytes = (size_t)w * h * ch;
    std::vector<uint8_t> packed;
    const uint8_t* src = frames[0].data;
    if (frames.size() > 1) {
      packed.resize(bytes * frames.size());
      for (size_t i = 0; i < frames.size(); ++i) {
        if (frames[i].width != w || frames[i].height != h || frames[i].channels != ch)
          throw std::runtime_error("ArucoDetector::detect: frames of one batch must have one size");
        std::copy(frames[i].data, frames[i].data + bytes, packed.begin() + i * bytes);
      }
      src = packed.data();
    }
    const int cap = 1024;
    std::vector<int32_t> n(frames.size()), ids(frames.size() * cap);
    std::vector<float> corners(frames.size() * cap * 8);
    if (arslam_detect_markers(det_, src, (int)frames.size(), w, h, ch, 0, &params_, cap, n.data(), ids.data(),
                              corners.data()) != ARSLAM_OK)
      throw std::runtime_error(std::string("arslam_detect_markers: ") + arslam_detector_last_error(det_));
    for (size_t i = 0; i < frames.size(); ++i) {
      auto& msg = out[i];
      msg.capture_uid = frames[i].capture_uid;
      msg.image_width = (uint32_t)w;
      msg.image_height = (uint32_t)h;
      msg.image_path = frames[i].image_path;
      msg.detector_types.emplace_back(detector_name_);
      msg.detections.resize(n[i]);
      for (int k = 0; k < n[i]; ++k) {
        auto& det = msg.detections[k];
        det.id = detector_name_ + '_' + std::to_string(ids[i * cap + k]);
        const float* c = &corners[(i * cap + k) * 8];
        for (int j = 0; j < 4; ++j) {       // from_cv_img: double arithmetic, stored as float32 Point32
          det.corners[j].x = (float)(c[2 * j] - 0.5 * w);
          det.corners[j].y = (float)(c[2 * j + 1] - 0.5 * h);
        }
      }
    }
    return out;
  }
  ar_slam_interfaces::msg::Detections detect(const Frame& frame) { return detect(std::vector<Frame>{frame})[0]; }

private:
  std::string detector_name_;
  arslam_detect_params params_;
  arslam_detector* det_ = nullptr;
};

}  // namespace ar_slam
