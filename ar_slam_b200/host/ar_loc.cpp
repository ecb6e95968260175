// ar_loc -- localise captures against a saved map (mirrors reference
// ar_slam/src/ar_loc.cpp:35-89): ar_loc map.yaml detect.yaml ...  -> localize.yaml
#include <fstream>
#include <iostream>

#include "ar_slam_solver.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "Need to provide a map .yaml and detection .yaml files" << std::endl;
    std::cerr << "Usage: ar_loc map.yaml [detect.yaml] ...\n"
                 "Description localizes captures from pre-processed detections against a map\n";
    return 1;
  }
  try {
    ArSlamSolver solver;
    const std::string map_fn = argv[1];
    std::cout << "Loading map " << map_fn << std::endl;
    solver.loadYaml(map_fn);
    const unsigned first_loc_cap_idx = solver.getNextCaptureIndex();
    for (int i = 2; i < argc; ++i) {
      const std::string fn = argv[i];
      if (!endswith(fn, ".yaml")) {
        std::cerr << "error loading image " << fn << " : image ingest (cv::aruco) is not part of this build, pass detections as .yaml" << std::endl;
        return 2;
      }
      solver.loadYaml(fn);
    }
    solver.localizeMany(first_loc_cap_idx);
    solver.printCameras();
    const std::string fn = "localize.yaml";
    std::cout << "Saving results to " << fn << std::endl;
    std::ofstream file(fn);
    solver.saveYaml(file);
  } catch (const std::exception& e) {
    std::cerr << "ar_loc: " << e.what() << std::endl;
    return 3;
  }
  return 0;
}
