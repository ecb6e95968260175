// host_selftest -- exercises the host facade without a GPU: yaml round trip, the data store,
// addDetections.  Used by tests/test_host_facade.py.
#include <iostream>
#include <sstream>

#include "ar_slam_solver.hpp"

int main(int argc, char** argv) {
  if (argc < 2 || (std::string(argv[1]) == "roundtrip" && argc < 3)) {
    std::cerr << "usage: host_selftest roundtrip in.yaml | detections" << std::endl;
    return 1;
  }
  const std::string mode = argv[1];
  try {
    if (mode == "roundtrip") {
      ArSlamSolver s;
      s.loadYaml(argv[2]);
      s.saveYaml(std::cout);
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
