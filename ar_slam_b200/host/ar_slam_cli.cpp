// ar_slam_cli -- map build from pre-processed detections, the ROS-free way to drive the
// solver (mirrors reference ar_slam/src/ar_slam_cli.cpp:33-81).  Image inputs need OpenCV's
// ArUco detector, which is upstream of the optimisation path and not built here.
#include <fstream>
#include <iostream>

#include "ar_slam_solver.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "Need to provide a .yaml of detections for processing" << std::endl;
    std::cerr << "Usage:  ar_slam_cli [fn1.yaml] [fn2.yaml] ...\n"
                 "Description run slam on pre-processed detections (yaml written by a previous run or by a detector)\n";
    return 1;
  }
  try {
    ArSlamSolver solver;
    for (int i = 1; i < argc; ++i) {
      const std::string fn = argv[i];
      if (!endswith(fn, ".yaml")) {
        std::cerr << "error loading image " << fn << " : image ingest (cv::aruco) is not part of this build, pass detections as .yaml" << std::endl;
        return 2;
      }
      solver.loadYaml(fn);
    }
    solver.solve();
    solver.printCameras();
    const std::string fn = "map.yaml";
    std::cout << "Saving results to " << fn << std::endl;
    std::ofstream file(fn);
    solver.saveYaml(file);
  } catch (const std::exception& e) {
    std::cerr << "ar_slam_cli: " << e.what() << std::endl;
    return 3;
  }
  return 0;
}
