// host_selftest -- exercises the host facade without a GPU: yaml round trip, the data store,
// addDetections.  Used by tests/test_host_facade.py.
#include <iostream>
#include <sstream>

#include <fstream>

#include "ar_slam_solver.hpp"
#include "aruco_detector.hpp"

int main(int argc, char** argv) {
  if (argc < 2 || ((std::string(argv[1]) == "roundtrip" || std::string(argv[1]) == "rosout") && argc < 3)) {
    std::cerr << "usage: host_selftest roundtrip in.yaml | detections | rosout map.yaml" << std::endl;
    return 1;
  }
  const std::string mode = argv[1];
  try {
    if (mode == "detect" && argc >= 4) {
      // host_selftest detect <dict> frame.pgm: the detector node's message for one grey netpbm frame (needs a GPU)
      std::ifstream in(argv[3], std::ios::binary);
      std::string magic;
      int w = 0, h = 0, maxv = 0;
      in >> magic >> w >> h >> maxv;
      in.get();
      if (magic != "P5" || maxv != 255) { std::cerr << "detect: binary 8-bit PGM only" << std::endl; return 1; }
      std::vector<uint8_t> px((size_t)w * h);
      in.read(reinterpret_cast<char*>(px.data()), (std::streamsize)px.size());
      ar_slam::ArucoDetector det(argv[2], 0, 1, w, h);
      det.params().min_corner_distance_rate = 0.1;
      ar_slam::Frame f;
      f.data = px.data(); f.width = w; f.height = h; f.channels = 1; f.capture_uid = "cap_test"; f.image_path = argv[3];
      const auto msg = det.detect(f);
      std::cout.precision(9);
      std::cout << "detections " << msg.capture_uid << " " << msg.image_width << " " << msg.image_height << " "
                << msg.detector_types.at(0) << " " << msg.detections.size() << "\n";
      for (const auto& d : msg.detections) {
        std::cout << d.id;
        for (const auto& c : d.corners) std::cout << " " << c.x << " " << c.y;
        std::cout << "\n";
      }
      return 0;
    }
    if (mode == "roundtrip") {
      ArSlamSolver s;
      s.loadYaml(argv[2]);
      s.saveYaml(std::cout);
      return 0;
    }
    if (mode == "rosout") {
      // the ROS outputs of the live component for a saved map, one record per line (17 significant digits)
      ArSlamSolver s;
      s.loadYaml(argv[2]);
      arslam_ros::Time stamp;
      stamp.sec = 12; stamp.nanosec = 345;
      std::cout.precision(17);
      for (const auto& t : s.getTransforms(stamp))
        std::cout << "tf " << t.header.frame_id << " " << t.child_frame_id << " " << t.header.stamp.sec << " " << t.header.stamp.nanosec << " "
                  << t.transform.translation.x << " " << t.transform.translation.y << " " << t.transform.translation.z << " "
                  << t.transform.rotation.x << " " << t.transform.rotation.y << " " << t.transform.rotation.z << " " << t.transform.rotation.w << "\n";
      const auto info = s.getCameraInfo();
      std::cout << "caminfo " << info.distortion_model << " " << info.d.size();
      for (double v : info.k) std::cout << " " << v;
      for (double v : info.r) std::cout << " " << v;
      for (double v : info.p) std::cout << " " << v;
      std::cout << "\n";
      std::vector<visualization_msgs::msg::Marker> markers;
      s.appendArucoMarkers(markers, stamp);
      for (const auto& m : markers)
        std::cout << "marker " << (m.header.frame_id.empty() ? "-" : m.header.frame_id) << " " << m.ns << " " << m.id << " " << m.type << " " << m.action << " "
                  << m.scale.x << " " << m.scale.y << " " << m.scale.z << " " << m.color.r << " " << m.color.g << " " << m.color.b << " " << m.color.a << " "
                  << (m.frame_locked ? 1 : 0) << "\n";
      return 0;
    }
    if (mode == "detections") {
      ArSlamSolver s;
      ar_slam_interfaces::msg::Detections d;
      d.capture_uid = "img1"; d.image_width = 1020; d.image_height = 768; d.image_path = "/tmp/img1.jpg";
      if (s.addDetections(d).has_value()) return 10;              // no detections -> nullopt
      ar_slam_interfaces::msg::Detection det;
      det.id = "aruco_4X4_50_7";
      for (int i = 0; i < 4; ++i) { det.corners[i].x = 0.1f * (i + 1); det.corners[i].y = -1.5f * i; }
      d.detections.push_back(det);
      auto h = s.addDetections(d);
      if (!h.has_value() || h->idx != 0) return 11;
      if (s.at(BlockHandle(0)).aruco_rect.corners[0].x != (double)0.1f) return 12;   // float32 widened, not re-rounded
      d.capture_uid = "img2"; d.image_width = 640;
      if (s.addDetections(d).has_value()) return 13;              // size mismatch is dropped
      d.image_width = 1020; d.capture_uid = "img1";
      try { s.addDetections(d); return 14; } catch (const std::runtime_error&) {}   // duplicate uid throws
      if (s.genUniqueCaptureUid().uid != "cap_1") return 15;
      if (filename_no_ext("../file.1.jpg") != "file.1" || filename_no_ext("/path/to/file.jpg") != "file") return 16;
      s.saveYaml(std::cout);
      return 0;
    }
  } catch (const std::exception& e) {
    std::cerr << "host_selftest: " << e.what() << std::endl;
    return 3;
  }
  return 1;
}
