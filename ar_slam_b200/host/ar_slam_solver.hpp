// ar_slam_solver.hpp -- host-side drop-in for the reference's solver facade.
//
// Mirrors the public surface of class ArSlamSolver
// (reference ar_slam/include/ar_slam/ar_slam_util.hpp:367-497) so that its
// callers -- ArSlam::detection_callback (ar_slam/src/ar_slam.cpp:114-156),
// ar_slam_cli (ar_slam_cli.cpp:58-78) and ar_loc (ar_loc.cpp:60-86) -- compile
// against it unchanged in what concerns the optimisation path.  The
// `ceres::Problem problem_` member is replaced by a handle of the CUDA library
// (include/ar_slam_b200.h); everything numeric happens on the GPU, the host
// keeps the data store, the schedules and map.yaml.
//
// Built without ROS2 / OpenCV / yaml-cpp / Ceres (none exist in this
// environment): the debug display (displayDebug) is
// outside the hot path and not provided; loadImages reads netpbm files instead
// of going through cv::imread and detects the markers on the GPU.
// The ROS output getters (getTransforms, getCameraInfo, appendArucoMarkers;
// reference ar_slam_util.cpp:1027-1162, called at ar_slam.cpp:133,145,154) ARE
// provided: they fill the real ROS2 messages when built with -DARSLAM_WITH_ROS
// and same-named plain structs otherwise (ros_msgs_lite.hpp).
#pragma once
#include <array>
#include <deque>
#include <iosfwd>
#include <optional>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/ar_slam_b200.h"
#include "detections_msg.hpp"
#include "ros_msgs_lite.hpp"

struct Point {
  double x = 0.0, y = 0.0;
};

struct CameraParams {
  std::array<double, 3> params{{3000.0, 0.0, 0.0}};  // focal, l1, l2; non-zero initial focal (hpp:69)
  std::optional<std::pair<int, int>> size;            // width, height
};

struct PoseParams {
  std::array<double, 6> params{{0, 0, 0, 0, 0, 0}};  // translation, angle-axis rotation
};

struct CaptureHandle { explicit CaptureHandle(unsigned i) : idx{i} {} bool operator==(const CaptureHandle& o) const { return o.idx == idx; } unsigned idx; };
struct ArucoHandle { explicit ArucoHandle(unsigned i) : idx{i} {} unsigned idx; };
struct BlockHandle { explicit BlockHandle(unsigned i) : idx{i} {} unsigned idx; };
struct CaptureUid { explicit CaptureUid(std::string u) : uid{std::move(u)} {} bool operator==(const CaptureUid& o) const { return o.uid == uid; } std::string uid; };
struct ArucoId { explicit ArucoId(std::string i) : id{std::move(i)} {} bool operator==(const ArucoId& o) const { return o.id == id; } std::string id; };

namespace std {
template <> struct hash<CaptureHandle> { size_t operator()(const CaptureHandle& h) const { return h.idx; } };
template <> struct hash<CaptureUid> { size_t operator()(const CaptureUid& u) const { return std::hash<std::string>{}(u.uid); } };
template <> struct hash<ArucoId> { size_t operator()(const ArucoId& a) const { return std::hash<std::string>{}(a.id); } };
}  // namespace std

struct ArucoRect {
  ArucoRect() = default;
  // Point32 corners are float32 and already centred; widened to double exactly as hpp:284-290
  explicit ArucoRect(const ar_slam_interfaces::msg::Detection& d) {
    for (unsigned i = 0; i < 4; ++i) { corners[i].x = d.corners[i].x; corners[i].y = d.corners[i].y; }
  }
  std::array<Point, 4> corners;
};

struct Capture {
  Capture(CaptureUid u, CaptureHandle h, std::string fn) : uid{std::move(u)}, handle{h}, img_fn{std::move(fn)} {}
  CaptureUid uid;
  CaptureHandle handle;
  std::string img_fn;
  std::vector<BlockHandle> blocks;
  std::optional<BlockHandle> init_block;
  PoseParams inv_pose;  // inverse pose: p_cam = R(w)(p_world + t)
  double* data() { return inv_pose.params.data(); }
  const double* data() const { return inv_pose.params.data(); }
};

struct Aruco {
  Aruco(ArucoId i, ArucoHandle h) : id{std::move(i)}, handle{h} {}
  ArucoId id;
  ArucoHandle handle;
  bool initialized = false;
  std::vector<BlockHandle> blocks;
  PoseParams pose;
  double* data() { return pose.params.data(); }
  const double* data() const { return pose.params.data(); }
};

struct Block {
  Block(BlockHandle h, const ArucoRect& r, CaptureHandle c, ArucoHandle a) : handle{h}, aruco_rect{r}, capture{c}, aruco{a} {}
  BlockHandle handle;
  ArucoRect aruco_rect;
  CaptureHandle capture;
  ArucoHandle aruco;
  bool added = false;
};

static constexpr double aruco_size = 0.0635;  // hpp:319

bool endswith(const std::string& str, const std::string& suffix);
std::string filename_no_ext(std::string filepath);

class ArSlamSolver {
public:
  ArSlamSolver();
  ~ArSlamSolver();
  ArSlamSolver(const ArSlamSolver&) = delete;
  ArSlamSolver& operator=(const ArSlamSolver&) = delete;

  // Image ingest (reference loadImages, ar_slam_util.cpp:247-286): marker detection on the GPU
  // (arslam_detect_markers, DICT_4X4_50, minCornerDistanceRate = 0.1), one capture per image, ids
  // "aruco_4X4_50_<n>".  Without OpenCV there is no cv::imread: the files are binary netpbm (P5 grey / P6 colour).
  void loadImages(const std::vector<std::string>& img_fns);
  void loadYaml(const std::string& fn);
  void saveYaml(std::ostream& output) const;
  void printCameras() const;
  void compareProjections() const;
  void solve();
  void solveIncremental();
  CaptureUid genUniqueCaptureUid() const;
  unsigned getNextCaptureIndex() const { return captures_.size(); }
  void localizeMany(unsigned first_loc_cap_idx);
  std::optional<CaptureHandle> addDetections(const ar_slam_interfaces::msg::Detections& detections);

  // ROS outputs of the live component (reference ar_slam_util.cpp:1027-1162)
  std::vector<geometry_msgs::msg::TransformStamped> getTransforms(const arslam_ros::Time& stamp) const;
  sensor_msgs::msg::CameraInfo getCameraInfo() const;
  void appendArucoMarkers(std::vector<visualization_msgs::msg::Marker>& markers, arslam_ros::Time stamp) const;

  bool& display_debug() { return display_debug_; }
  double& display_wait_duration() { return display_wait_duration_; }

  Capture& at(CaptureHandle h) { return captures_[h.idx]; }
  const Capture& at(CaptureHandle h) const { return captures_[h.idx]; }
  Aruco& at(ArucoHandle h) { return arucos_[h.idx]; }
  const Aruco& at(ArucoHandle h) const { return arucos_[h.idx]; }
  Block& at(BlockHandle h) { return blocks_[h.idx]; }
  const Block& at(BlockHandle h) const { return blocks_[h.idx]; }

  // additions (not in the reference)
  // Schedules with the parameters resident on the GPU (default): between the optimize() calls of one
  // solve() / solveIncremental() nothing but the new capture's blocks crosses the bus -- new captures and
  // tags are seeded on the device (arslam_seed_captures / arslam_seed_tags) from the previous solve's result,
  // and the poses come back once, when the schedule returns.  false: every optimize() uploads and downloads
  // all parameters, like the reference's Ceres calls touch the caller's arrays.
  bool& device_resident_schedule() { return device_resident_; }
  // The reference's own TODO (ar_slam_util.cpp:810): add this many captures to the problem per optimize()
  // call in solve().  1 reproduces the reference's schedule; larger values trade its exact trajectory for
  // fewer (quadratically growing) solves.
  unsigned& captures_per_solve() { return captures_per_solve_; }
  // creates the GPU handle now (CUDA context, kernel attributes: about a second in a fresh process) instead of
  // inside the first optimize(); a long-lived node pays this once at start-up
  void prepareDevice() { handle(); }
  // the summary Ceres would have returned, and the solver options
  const arslam_summary& lastSummary() const { return last_summary_; }
  const std::vector<arslam_summary>& summaries() const { return summaries_; }
  arslam_options& options() { return options_; }
  const CameraParams& camera() const { return camera_; }
  size_t numCaptures() const { return captures_.size(); }
  size_t numArucos() const { return arucos_.size(); }
  size_t numBlocks() const { return blocks_.size(); }

protected:
  Capture& addCapture(CaptureUid cap_uid, std::string fn);
  Aruco& addAruco(ArucoId ar_id);
  Aruco& getOrAddAruco(const ArucoId& ar_id);
  Block& addBlock(const ArucoRect& aruco_rect, CaptureHandle capture_handle, ArucoHandle aruco_handle);
  void addConnectedCaptures(const Capture& base_capture, std::deque<CaptureHandle>& open_captures);
  void solveCapture(Capture& capture, std::optional<BlockHandle> init_block_handle);
  void addCaptureBlocksToProblem(Capture& capture);
  void optimize(const Capture& capture);
  void resetProblem();
  arslam_solver* handle();
  // seeds go to the host arrays, or -- while the parameters live on the GPU -- are queued for the device
  void seedCapture(Capture& capture, const Block& from_block);
  void seedAruco(Aruco& aruco, const Capture& from_capture, const Block& block);
  void syncFromDevice();   // downloads the parameters when the host copies are stale
  struct PendingSeed { bool is_capture; int32_t target, source; double rect[8]; };
  std::vector<PendingSeed> pending_seeds_;
  bool device_resident_ = true;   // option
  bool device_params_ = false;    // the GPU holds the current parameters of the device_captures_ / device_arucos_ poses
  bool host_stale_ = false;       // ... and the host copies are older than those
  unsigned captures_per_solve_ = 1;

  arslam_solver* gpu_ = nullptr;              // replaces `ceres::Problem problem_` (hpp:473)
  arslam_options options_;
  std::vector<BlockHandle> problem_blocks_;   // residual blocks in AddResidualBlock order
  std::vector<BlockHandle> device_blocks_;    // the blocks resident on the GPU, in upload order (incremental uploads)
  size_t device_captures_ = 0, device_arucos_ = 0;
  arslam_summary last_summary_{};
  std::vector<arslam_summary> summaries_;

  CameraParams camera_;
  std::deque<Capture> captures_;
  std::deque<Aruco> arucos_;
  std::vector<Block> blocks_;
  std::unordered_map<CaptureUid, CaptureHandle> capture_map_;
  std::unordered_map<ArucoId, ArucoHandle> aruco_map_;
  std::unordered_set<CaptureHandle> unsolved_captures_;
  bool display_debug_ = false;   // no GUI here; the reference's node default is false as well (ar_slam.cpp:74)
  double display_wait_duration_ = 0.0;
};
