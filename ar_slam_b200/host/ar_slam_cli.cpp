// ar_slam_cli -- map build from pre-processed detections, the ROS-free way to drive the
// solver (mirrors reference ar_slam/src/ar_slam_cli.cpp:33-81).  Image inputs need OpenCV's
// ArUco detector, which is upstream of the optimisation path and not built here.
// Options that the reference does not have (they must precede the file names):
//   --host-params           every optimize() moves all parameters host <-> device (the reference's data flow)
//                           instead of keeping them on the GPU for the whole schedule
//   --captures-per-solve K  add K captures per optimize() call (the reference's TODO at ar_slam_util.cpp:810)
//   --quiet                 no per-capture progress lines
//   --output FILE           where the map goes (default map.yaml)
#include <fstream>
#include <chrono>
#include <cstdlib>
#include <iostream>
#include <sstream>

#include "ar_slam_solver.hpp"

int main(int argc, char** argv) {
  if (argc < 2) {
    std::cerr << "Need to provide a .yaml of detections for processing" << std::endl;
    std::cerr << "Usage:  ar_slam_cli [fn1.yaml | img1.ppm] [fn2.yaml | img2.pgm] ...\n"
                 "Description run slam on pre-processed detections (yaml) or on images (binary netpbm; the markers are "
                 "detected on the GPU)\n";
    return 1;
  }
  try {
    ArSlamSolver solver;
    std::string out_fn = "map.yaml";
    bool quiet = false, solve_log = false;
    int first_file = 1;
    while (first_file < argc && std::string(argv[first_file]).rfind("--", 0) == 0) {
      const std::string opt = argv[first_file++];
      if (opt == "--host-params") solver.device_resident_schedule() = false;
      else if (opt == "--quiet") quiet = true;
      else if (opt == "--solve-log") solve_log = true;
      else if (opt == "--captures-per-solve" && first_file < argc) solver.captures_per_solve() = (unsigned)std::atoi(argv[first_file++]);
      else if (opt == "--output" && first_file < argc) out_fn = argv[first_file++];
      else { std::cerr << "unknown option " << opt << std::endl; return 1; }
    }
    if (first_file >= argc) { std::cerr << "Need to provide a .yaml of detections for processing" << std::endl; return 1; }
    std::streambuf* cout_buf = std::cout.rdbuf();
    std::ostringstream sink;
    if (quiet) std::cout.rdbuf(sink.rdbuf());
    // ar_slam_cli.cpp:58-70 of the reference: .yaml files are detections, everything else is an image
    std::vector<std::string> img_fns;
    for (int i = first_file; i < argc; ++i) {
      const std::string fn = argv[i];
      if (endswith(fn, ".yaml")) solver.loadYaml(fn);
      else img_fns.push_back(fn);
    }
    if (!img_fns.empty()) solver.loadImages(img_fns);   // netpbm frames, markers detected on the GPU
    solver.prepareDevice();  // context creation is not part of the schedule
    const auto t0 = std::chrono::steady_clock::now();
    solver.solve();
    const double schedule_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (quiet) std::cout.rdbuf(cout_buf);
    if (solve_log) {
      std::cout.precision(12);
      int i = 0;
      for (const arslam_summary& sm : solver.summaries())
        std::cout << "solve " << i++ << " iterations " << sm.iterations << " initial_cost " << sm.initial_cost << " final_cost " << sm.final_cost
                  << " termination " << sm.termination << " reason " << sm.reason << " elim " << sm.eliminated_side << " lin " << sm.linear_solver
                  << " reduced_dim " << sm.reduced_dim << std::endl;
      std::cout.precision(6);
    }
    solver.printCameras();
    {
      // what the schedule cost: optimize() calls, LM iterations, device and host time of the solves
      long long its = 0; double ms = 0.0, dev_ms = 0.0;
      for (const arslam_summary& sm : solver.summaries()) { its += sm.iterations; ms += sm.total_ms; dev_ms += sm.eval_ms + sm.linsolve_ms; }
      std::cout << "schedule: " << solver.summaries().size() << " solves, " << its << " LM iterations, " << ms << " ms in arslam_solve ("
                << dev_ms << " ms on the device), " << schedule_ms << " ms for the whole solve() schedule, final cost "
                << solver.lastSummary().final_cost << std::endl;
    }
    const std::string fn = out_fn;
    std::cout << "Saving results to " << fn << std::endl;
    std::ofstream file(fn);
    solver.saveYaml(file);
  } catch (const std::exception& e) {
    std::cerr << "ar_slam_cli: " << e.what() << std::endl;
    return 3;
  }
  return 0;
}
